// TEST INFRASTRUCTURE ONLY (never linked into the product).
//
// ref_dump: a small driver over the UNMODIFIED reference library's public API
// (PopSift::enqueue / SiftJob::get, /root/reference/src/popsift/popsift.h:105-317),
// compiled by oracle/build_ref.sh against oracle/_ref/libpopsift_ref.so.
// It is our own code: it only calls the reference, it contains none of it.
//
//  ref_dump -i frame.pgm [-i more.pgm ...] -o features.bin [config flags]
//           [--log]               reference LogMode::All: raw plane dumps into cwd
//                                 (dir-octave-dump/, dir-dog-dump/, sift_octave.cu:111-188)
//           [--bench STEPS WARMUP] time STEPS passes over all frames, print JSON
//           [--float-mode]        PopSift::FloatImages, pixels = u8 / 256 (reference main.cpp:234)
//           [--filter-max-extrema N --filter-grid G --filter-sort up|down|random]   grid filter
//           [--desc-mode loop|iloop|grid|igrid|notile] [--direct-scaling]
//           [--match]             Config::MatchingMode on exactly two inputs: SiftJob::getDev for both,
//                                 device results copied back into features.bin.0/.1 (+ "PSR1" reverse map),
//                                 then FeaturesDev::match (features.cu:282-304), whose device printf lines
//                                 ("accept feat ..." / "reject feat ...") go to stdout
//           [--repeat N]          run the whole frame list N times, outputs suffixed .rK (run-to-run jitter)
//
// features.bin layout (little endian):
//   char magic[4]="PSF1"; int32 n_feat; int32 n_desc;
//   n_feat x { int32 octave; float x,y,sigma; int32 num_ori; float ori[4]; int32 desc_idx[4]; }
//   n_desc x float[128]
#include <popsift/popsift.h>
#include <popsift/features.h>
#include <popsift/sift_conf.h>
#include <popsift/sift_extremum.h>
#include <cuda_runtime.h>

#include <chrono>
#include <unistd.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <string>
#include <vector>

struct Frame { int w = 0, h = 0; std::vector<unsigned char> px; };

static bool read_pgm(const std::string& fn, Frame& f)
{
    std::ifstream in(fn, std::ios::binary);
    if (!in) return false;
    std::string magic; in >> magic;
    if (magic != "P5") return false;
    auto skip = [&]() { in >> std::ws; while (in.peek() == '#') { std::string l; std::getline(in, l); in >> std::ws; } };
    int maxv;
    skip(); in >> f.w; skip(); in >> f.h; skip(); in >> maxv;
    in.get();
    if (maxv != 255) return false;
    f.px.resize(size_t(f.w) * f.h);
    in.read((char*)f.px.data(), f.px.size());
    return bool(in);
}

static void write_features(const std::string& fn, popsift::FeaturesHost* fh)
{
    FILE* fp = fopen(fn.c_str(), "wb");
    if (!fp) { perror("fopen"); exit(2); }
    int nf = fh->getFeatureCount(), nd = fh->getDescriptorCount();
    fwrite("PSF1", 1, 4, fp); fwrite(&nf, 4, 1, fp); fwrite(&nd, 4, 1, fp);
    popsift::Feature* F = fh->getFeatures();
    popsift::Descriptor* D = fh->getDescriptors();
    for (int i = 0; i < nf; i++) {
        int idx[4];
        for (int k = 0; k < 4; k++) idx[k] = (k < F[i].num_ori && F[i].desc[k]) ? int(F[i].desc[k] - D) : -1;
        fwrite(&F[i].debug_octave, 4, 1, fp);
        fwrite(&F[i].xpos, 4, 1, fp); fwrite(&F[i].ypos, 4, 1, fp); fwrite(&F[i].sigma, 4, 1, fp);
        fwrite(&F[i].num_ori, 4, 1, fp); fwrite(F[i].orientation, 4, 4, fp); fwrite(idx, 4, 4, fp);
    }
    fwrite(D, sizeof(float) * 128, nd, fp);
    fclose(fp);
}

// device-resident results (Config::MatchingMode) copied back; Feature::desc[] are DEVICE pointers into the
// descriptor array and become indices; the reverse map follows as "PSR1" n_desc x int32
static void write_dev_features(const std::string& fn, popsift::FeaturesDev* fd)
{
    const int nf = fd->getFeatureCount(), nd = fd->getDescriptorCount();
    std::vector<popsift::Feature> F(nf);
    std::vector<popsift::Descriptor> D(nd);
    std::vector<int> R(nd);
    cudaMemcpy(F.data(), fd->getFeatures(), sizeof(popsift::Feature) * nf, cudaMemcpyDeviceToHost);
    cudaMemcpy(D.data(), fd->getDescriptors(), sizeof(popsift::Descriptor) * nd, cudaMemcpyDeviceToHost);
    cudaMemcpy(R.data(), fd->getReverseMap(), sizeof(int) * nd, cudaMemcpyDeviceToHost);
    FILE* fp = fopen(fn.c_str(), "wb");
    if (!fp) { perror("fopen"); exit(2); }
    fwrite("PSF1", 1, 4, fp); fwrite(&nf, 4, 1, fp); fwrite(&nd, 4, 1, fp);
    popsift::Descriptor* base = fd->getDescriptors();
    for (int i = 0; i < nf; i++) {
        int idx[4];
        for (int k = 0; k < 4; k++) idx[k] = (k < F[i].num_ori && F[i].desc[k]) ? int(F[i].desc[k] - base) : -1;
        fwrite(&F[i].debug_octave, 4, 1, fp);
        fwrite(&F[i].xpos, 4, 1, fp); fwrite(&F[i].ypos, 4, 1, fp); fwrite(&F[i].sigma, 4, 1, fp);
        fwrite(&F[i].num_ori, 4, 1, fp); fwrite(F[i].orientation, 4, 4, fp); fwrite(idx, 4, 4, fp);
    }
    fwrite(D.data(), sizeof(float) * 128, nd, fp);
    fwrite("PSR1", 1, 4, fp);
    fwrite(R.data(), 4, nd, fp);
    fclose(fp);
}

int main(int argc, char** argv)
{
    std::vector<std::string> inputs;
    std::string out;
    bool log = false, float_mode = false, do_match = false, direct_scaling = false;
    int bench_steps = 0, bench_warm = 0, device = 0, repeat = 1;
    int filter_max = -1, filter_grid = -1;
    std::string filter_sort = "", desc_mode = "", go_file = "";
    std::string mode = "popsift", norm = "", gauss = "";
    float downsampling = 1e9f, sigma = -1, threshold = -1, edge = -1, iblur = -1;
    int octaves = -2, levels = -1, norm_multi = -1000;

    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto nxt = [&]() -> const char* { if (i + 1 >= argc) { fprintf(stderr, "missing value for %s\n", a.c_str()); exit(2); } return argv[++i]; };
        if (a == "-i") inputs.push_back(nxt());
        else if (a == "-o") out = nxt();
        else if (a == "--log") log = true;
        else if (a == "--mode") mode = nxt();
        else if (a == "--norm") norm = nxt();
        else if (a == "--gauss-mode") gauss = nxt();
        else if (a == "--downsampling") downsampling = atof(nxt());
        else if (a == "--octaves") octaves = atoi(nxt());
        else if (a == "--levels") levels = atoi(nxt());
        else if (a == "--sigma") sigma = atof(nxt());
        else if (a == "--threshold") threshold = atof(nxt());
        else if (a == "--edge-limit") edge = atof(nxt());
        else if (a == "--initial-blur") iblur = atof(nxt());
        else if (a == "--norm-multi") norm_multi = atoi(nxt());
        else if (a == "--device") device = atoi(nxt());
        else if (a == "--float-mode") float_mode = true;
        else if (a == "--match") do_match = true;
        else if (a == "--direct-scaling") direct_scaling = true;
        else if (a == "--repeat") repeat = atoi(nxt());
        else if (a == "--filter-max-extrema") filter_max = atoi(nxt());
        else if (a == "--filter-grid") filter_grid = atoi(nxt());
        else if (a == "--filter-sort") filter_sort = nxt();
        else if (a == "--desc-mode") desc_mode = nxt();
        else if (a == "--go-file") go_file = nxt();     // --bench: after the warm-up wait until this file exists
        else if (a == "--bench") { bench_steps = atoi(nxt()); bench_warm = atoi(nxt()); }
        else { fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
    if (inputs.empty()) { fprintf(stderr, "no input\n"); return 2; }

    cudaSetDevice(device);
    popsift::Config cfg;   // needs a CUDA device (sift_conf.cu:46-50)
    if (mode == "vlfeat") cfg.setMode(popsift::Config::VLFeat);
    else if (mode == "opencv") cfg.setMode(popsift::Config::OpenCV);
    else cfg.setMode(popsift::Config::PopSift);
    if (!gauss.empty()) cfg.setGaussMode(gauss);
    if (norm == "classic") cfg.setNormMode(popsift::Config::Classic);
    else if (norm == "rootsift") cfg.setNormMode(popsift::Config::RootSift);
    if (downsampling < 1e8f) cfg.setDownsampling(downsampling);
    if (octaves != -2) cfg.setOctaves(octaves);
    if (levels > 0) cfg.setLevels(levels);
    if (sigma > 0) cfg.setSigma(sigma);
    if (threshold >= 0) cfg.setThreshold(threshold);
    if (edge >= 0) cfg.setEdgeLimit(edge);
    if (iblur >= 0) cfg.setInitialBlur(iblur);
    if (norm_multi > -1000) cfg.setNormalizationMultiplier(norm_multi);
    if (log) cfg.setLogMode(popsift::Config::All);
    if (filter_max >= 0) cfg.setFilterMaxExtrema(filter_max);
    if (filter_grid >= 0) cfg.setFilterGridSize(filter_grid);
    if (!filter_sort.empty()) cfg.setFilterSorting(filter_sort);
    if (!desc_mode.empty()) cfg.setDescMode(desc_mode);
    if (direct_scaling) cfg.setScalingMode(popsift::Config::ScaleDirect);

    std::vector<Frame> frames(inputs.size());
    for (size_t k = 0; k < inputs.size(); k++)
        if (!read_pgm(inputs[k], frames[k])) { fprintf(stderr, "cannot read %s\n", inputs[k].c_str()); return 2; }

    if (do_match) {
        if (frames.size() != 2) { fprintf(stderr, "--match needs exactly two inputs\n"); return 2; }
        PopSift msift(cfg, popsift::Config::MatchingMode, PopSift::ByteImages, device);
        popsift::FeaturesDev* fd[2];
        for (int k = 0; k < 2; k++) {
            SiftJob* j = msift.enqueue(frames[k].w, frames[k].h, frames[k].px.data());
            fd[k] = j->getDev();
            delete j;
            fprintf(stderr, "ref_dump: %s -> %d features, %d descriptors (device)\n", inputs[k].c_str(),
                    fd[k]->getFeatureCount(), fd[k]->getDescriptorCount());
            if (!out.empty()) write_dev_features(out + "." + std::to_string(k), fd[k]);
        }
        fflush(stdout);
        fd[0]->match(fd[1]);          // device printf -> stdout
        cudaDeviceSynchronize();
        fflush(stdout);
        delete fd[0]; delete fd[1];
        msift.uninit();
        return 0;
    }

    PopSift sift(cfg, popsift::Config::ExtractingMode, float_mode ? PopSift::FloatImages : PopSift::ByteImages, device);
    std::vector<std::vector<float>> fframes;
    if (float_mode) {
        fframes.resize(frames.size());
        for (size_t k = 0; k < frames.size(); k++) {
            fframes[k].resize(frames[k].px.size());
            for (size_t i = 0; i < frames[k].px.size(); i++) fframes[k][i] = float(frames[k].px[i]) / 256.0f;
        }
    }
    auto enqueue = [&](size_t k) -> SiftJob* {
        return float_mode ? sift.enqueue(frames[k].w, frames[k].h, fframes[k].data())
                          : sift.enqueue(frames[k].w, frames[k].h, frames[k].px.data());
    };

    if (bench_steps > 0) {
        // one "step" = every frame enqueued, every result fetched (host buffers in, host features out)
        size_t nfeat = 0, ndesc = 0;
        auto pass = [&]() {
            std::vector<SiftJob*> jobs;
            for (size_t k = 0; k < frames.size(); k++) jobs.push_back(enqueue(k));
            nfeat = ndesc = 0;
            for (auto* j : jobs) {
                if (!j) continue;
                popsift::FeaturesHost* fh = j->get();
                nfeat += fh->getFeatureCount(); ndesc += fh->getDescriptorCount();
                delete fh; delete j;
            }
        };
        for (int w = 0; w < bench_warm; w++) pass();
        cudaDeviceSynchronize();
        if (!go_file.empty()) {
            printf("ready\n"); fflush(stdout);
            while (!std::ifstream(go_file).good()) usleep(200);
        }
        auto t0 = std::chrono::steady_clock::now();
        for (int s = 0; s < bench_steps; s++) pass();
        cudaDeviceSynchronize();
        auto t1 = std::chrono::steady_clock::now();
        double ms = std::chrono::duration<double, std::milli>(t1 - t0).count();
        double px = 0; for (auto& f : frames) px += double(f.w) * f.h;
        printf("{\"ref_bench\": true, \"steps\": %d, \"warmup\": %d, \"frames_per_step\": %zu, "
               "\"pixels_per_step\": %.0f, \"total_ms\": %.4f, \"ms_per_step\": %.4f, "
               "\"mpix_per_s\": %.4f, \"features_last_step\": %zu, \"descriptors_last_step\": %zu}\n",
               bench_steps, bench_warm, frames.size(), px, ms, ms / bench_steps,
               px * bench_steps / (ms * 1e-3) / 1e6, nfeat, ndesc);
        fflush(stdout);
    } else {
        for (int rep = 0; rep < repeat; rep++)
        for (size_t k = 0; k < frames.size(); k++) {
            SiftJob* j = enqueue(k);
            if (!j) { fprintf(stderr, "enqueue failed\n"); return 3; }
            popsift::FeaturesHost* fh = j->get();
            fprintf(stderr, "ref_dump: %s -> %d features, %d descriptors\n", inputs[k].c_str(),
                    fh->getFeatureCount(), fh->getDescriptorCount());
            if (!out.empty()) {
                std::string fn = out;
                if (frames.size() > 1) fn += "." + std::to_string(k);
                if (repeat > 1) fn += ".r" + std::to_string(rep);
                write_features(fn, fh);
            }
            delete fh; delete j;
        }
    }
    sift.uninit();
    return 0;
}
