// Exercises the C++ drop-in API exactly as a libpopsift caller would (reference usage:
// src/application/main.cpp:172-264): Config -> PopSift -> enqueue -> SiftJob::get -> Features iteration.
// Usage: api_check W H seed-file.raw [vlfeat|popsift] [classic|rootsift] [nframes]
// Prints one line per frame: "<n_features> <n_descriptors> <checksum>".
#include <popsift/features.h>
#include <popsift/popsift.h>
#include <popsift/sift_conf.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

int main(int argc, char** argv)
{
    if (argc < 4) { std::fprintf(stderr, "usage: api_check W H file.raw [mode] [norm] [n]\n"); return 2; }
    const int w = std::atoi(argv[1]), h = std::atoi(argv[2]);
    std::vector<unsigned char> img((size_t)w * h);
    std::ifstream in(argv[3], std::ios::binary);
    in.read(reinterpret_cast<char*>(img.data()), img.size());
    if (!in) { std::fprintf(stderr, "short read\n"); return 2; }
    const std::string mode = argc > 4 ? argv[4] : "popsift", norm = argc > 5 ? argv[5] : "rootsift";
    const int n = argc > 6 ? std::atoi(argv[6]) : 1;

    popart::Config config;                       // the README's spelling of the namespace must compile too
    if (mode == "vlfeat") config.setMode(popsift::Config::VLFeat);
    config.setNormMode(norm == "classic" ? popsift::Config::Classic : popsift::Config::RootSift);
    popsift::Config same = config;
    if (!(same == config)) return 3;

    PopSift sift(config, popsift::Config::ExtractingMode, PopSift::ByteImages, 0);
    std::vector<SiftJob*> jobs;
    for (int i = 0; i < n; ++i) jobs.push_back(sift.enqueue(w, h, img.data()));
    for (SiftJob* job : jobs) {
        popsift::Features* f = job->get();
        double sum = 0.0;
        int nd = 0;
        for (const popsift::Feature& ft : *f)
            for (int o = 0; o < ft.num_ori; ++o) {
                ++nd;
                for (int k = 0; k < 128; ++k) sum += ft.desc[o]->features[k];
            }
        if (nd != f->getDescriptorCount()) return 4;
        std::printf("%d %d %.6f\n", f->getFeatureCount(), f->getDescriptorCount(), sum);
        delete f;
        delete job;
    }
    // wrong image mode must throw on the caller's thread (reference popsift.cpp:247-253)
    bool threw = false;
    try { std::vector<float> fi((size_t)w * h, 0.f); sift.enqueue(w, h, fi.data()); } catch (const std::runtime_error&) { threw = true; }
    if (!threw) return 5;
    sift.uninit();

    // Config::MatchingMode: results stay on the device (reference popsift.h:79-86, features.h:104-122)
    {
        PopSift msift(config, popsift::Config::MatchingMode, PopSift::ByteImages, 0);
        SiftJob* job = msift.enqueue(w, h, img.data());
        popsift::FeaturesDev* fd = job->getDev();
        if (!fd) return 6;
        if (fd->getFeatureCount() <= 0 || fd->getDescriptorCount() < fd->getFeatureCount()) return 7;
        if (!fd->getFeatures() || !fd->getDescriptors() || !fd->getReverseMap()) return 8;
        // the matcher (reference features.cu:282-304): every descriptor's nearest neighbour in its own set is itself
        // (or an identical earlier descriptor), at distance 0
        const std::vector<int> m = fd->matchIndices(fd);
        if ((int)m.size() != 3 * fd->getDescriptorCount()) return 9;
        int self = 0;
        for (int i = 0; i < fd->getDescriptorCount(); ++i) {
            if (m[3 * i] < 0 || m[3 * i] >= fd->getDescriptorCount() || m[3 * i + 1] < 0 || m[3 * i + 1] >= fd->getDescriptorCount()) return 10;
            if (m[3 * i] == i) ++self;
        }
        if (self < fd->getDescriptorCount() * 9 / 10) return 11;
        std::printf("dev %d %d\n", fd->getFeatureCount(), fd->getDescriptorCount());
        delete fd;
        delete job;
        msift.uninit();
    }
    return 0;
}
