// End-to-end timing through the C++ drop-in API, the way a libpopsift caller drives it and the way
// oracle/ref_driver.cpp --bench drives the unmodified reference (reference usage: src/application/main.cpp:172-264):
// pageable std::vector frames -> PopSift::enqueue -> SiftJob::get -> delete.  One "step" = every frame enqueued,
// every result fetched; host -> device copies of all frames and device -> host copies of all results are inside
// the timed region (wall clock around cudaDeviceSynchronize-free API calls: get() returns host data).
//
//   api_bench [--octaves N] [--levels N] [--mode popsift|vlfeat|opencv] [--norm rootsift|classic] [--device D]
//             [--slots S] [--go-file PATH] --bench STEPS WARMUP -i f0.pgm [-i f1.pgm ...]
// --go-file: after the warm-up, wait until PATH exists before the timed steps (lets several ranks start together).
// Prints one JSON line (same keys as ref_dump --bench).
#include <popsift/features.h>
#include <popsift/popsift.h>
#include <popsift/sift_conf.h>

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <thread>
#include <vector>

struct Frame { int w = 0, h = 0; std::vector<unsigned char> px; };

static bool read_pgm(const std::string& fn, Frame& f)
{
    std::ifstream in(fn, std::ios::binary);
    if (!in) return false;
    std::string magic; in >> magic;
    if (magic != "P5") return false;
    int maxv = 0;
    in >> f.w >> f.h >> maxv;
    in.get();
    if (maxv != 255 || f.w < 1 || f.h < 1) return false;
    f.px.resize((size_t)f.w * f.h);
    in.read(reinterpret_cast<char*>(f.px.data()), (std::streamsize)f.px.size());
    return (bool)in;
}

int main(int argc, char** argv)
{
    std::vector<std::string> inputs;
    std::string mode = "popsift", norm = "rootsift", go_file;
    int octaves = -2, levels = -1, device = 0, slots = 4, steps = 0, warm = 0;
    for (int i = 1; i < argc; ++i) {
        const std::string a = argv[i];
        auto nxt = [&]() -> const char* { if (i + 1 >= argc) { std::fprintf(stderr, "missing value for %s\n", a.c_str()); std::exit(2); } return argv[++i]; };
        if (a == "-i") inputs.push_back(nxt());
        else if (a == "--octaves") octaves = std::atoi(nxt());
        else if (a == "--levels") levels = std::atoi(nxt());
        else if (a == "--mode") mode = nxt();
        else if (a == "--norm") norm = nxt();
        else if (a == "--device") device = std::atoi(nxt());
        else if (a == "--slots") slots = std::atoi(nxt());
        else if (a == "--go-file") go_file = nxt();
        else if (a == "--bench") { steps = std::atoi(nxt()); warm = std::atoi(nxt()); }
        else { std::fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
    if (inputs.empty() || steps < 1) { std::fprintf(stderr, "usage: api_bench --bench STEPS WARMUP -i f.pgm ...\n"); return 2; }

    popsift::Config cfg;
    if (mode == "vlfeat") cfg.setMode(popsift::Config::VLFeat);
    else if (mode == "opencv") cfg.setMode(popsift::Config::OpenCV);
    cfg.setNormMode(norm == "classic" ? popsift::Config::Classic : popsift::Config::RootSift);
    if (octaves != -2) cfg.setOctaves(octaves);
    if (levels > 0) cfg.setLevels(levels);

    std::vector<Frame> frames(inputs.size());
    for (size_t k = 0; k < inputs.size(); ++k)
        if (!read_pgm(inputs[k], frames[k])) { std::fprintf(stderr, "cannot read %s\n", inputs[k].c_str()); return 2; }

    PopSift sift(cfg, popsift::Config::ExtractingMode, PopSift::ByteImages, device);
    sift.setSlots(slots);
    size_t nfeat = 0, ndesc = 0;
    auto pass = [&]() {
        std::vector<SiftJob*> jobs;
        jobs.reserve(frames.size());
        for (auto& f : frames) jobs.push_back(sift.enqueue(f.w, f.h, f.px.data()));
        nfeat = ndesc = 0;
        for (SiftJob* j : jobs) {
            if (!j) continue;
            popsift::FeaturesHost* fh = j->get();
            nfeat += (size_t)fh->getFeatureCount(); ndesc += (size_t)fh->getDescriptorCount();
            delete fh; delete j;
        }
    };
    for (int w = 0; w < warm; ++w) pass();
    if (!go_file.empty()) {
        std::printf("ready\n"); std::fflush(stdout);
        while (!std::ifstream(go_file).good()) std::this_thread::sleep_for(std::chrono::microseconds(200));
    }
    const auto t0 = std::chrono::steady_clock::now();
    for (int s = 0; s < steps; ++s) pass();
    const auto t1 = std::chrono::steady_clock::now();
    const double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
    double px = 0; for (auto& f : frames) px += (double)f.w * f.h;
    std::printf("{\"api_bench\": true, \"steps\": %d, \"warmup\": %d, \"frames_per_step\": %zu, \"pixels_per_step\": %.0f, "
                "\"total_ms\": %.4f, \"ms_per_step\": %.4f, \"mpix_per_s\": %.4f, \"features_last_step\": %zu, "
                "\"descriptors_last_step\": %zu, \"slots\": %d}\n",
                steps, warm, frames.size(), px, ms, ms / steps, px * steps / (ms * 1e-3) / 1e6, nfeat, ndesc, slots);
    std::fflush(stdout);
    sift.uninit();
    return 0;
}
