// popsift-demo for the Blackwell-native drop-in: same command line, same output file and the same
// console lines as the reference's demo (reference src/application/main.cpp:49-330), with a
// hand-rolled option parser (no Boost) and std::filesystem for directory recursion.
#include <popsift/common/device_prop.h>
#include <popsift/features.h>
#include <popsift/popsift.h>
#include <popsift/sift_conf.h>
#include <popsift/sift_config.h>
#include <popsift/version.hpp>

#include "pgmread.h"

#include <cstdlib>
#include <filesystem>
#include <fstream>
#include <functional>
#include <iostream>
#include <list>
#include <map>
#include <queue>
#include <string>

using namespace std;
namespace fs = std::filesystem;

static bool print_dev_info = false, print_time_info = false, write_as_uchar = false, dont_write = false,
            pgmread_loading = false, float_mode = false;

struct Opt { bool has_arg; function<void(const string&)> fn; string help; };

static void usage(const map<string, Opt>& opts)
{
    cout << "Allowed options:" << endl;
    for (auto& kv : opts) cout << "  --" << kv.first << (kv.second.has_arg ? " arg" : "") << "\t" << kv.second.help << endl;
}

static void parseargs(int argc, char** argv, popsift::Config& config, string& inputFile)
{
    map<string, Opt> o;
    auto flag = [&](const string& n, function<void()> f, const string& h) { o[n] = {false, [f](const string&) { f(); }, h}; };
    auto val = [&](const string& n, function<void(const string&)> f, const string& h) { o[n] = {true, f, h}; };
    flag("help", [&] { usage(o); exit(EXIT_SUCCESS); }, "Print usage");
    flag("verbose", [&] { config.setVerbose(); }, "");
    flag("log", [&] { config.setLogMode(popsift::Config::All); }, "Write debugging files");
    val("input-file", [&](const string& s) { inputFile = s; }, "Input file");
    val("octaves", [&](const string& s) { config.octaves = atoi(s.c_str()); }, "Number of octaves");
    val("levels", [&](const string& s) { config.levels = atoi(s.c_str()); }, "Number of levels per octave");
    val("sigma", [&](const string& s) { config.setSigma((float)atof(s.c_str())); }, "Initial sigma value");
    val("threshold", [&](const string& s) { config.setThreshold((float)atof(s.c_str())); }, "Contrast threshold");
    val("edge-threshold", [&](const string& s) { config.setEdgeLimit((float)atof(s.c_str())); }, "On-edge threshold");
    val("edge-limit", [&](const string& s) { config.setEdgeLimit((float)atof(s.c_str())); }, "On-edge threshold");
    val("downsampling", [&](const string& s) { config.setDownsampling((float)atof(s.c_str())); }, "Downscale width and height of input by 2^N");
    val("initial-blur", [&](const string& s) { config.setInitialBlur((float)atof(s.c_str())); }, "Assume initial blur, subtract when blurring first time");
    val("gauss-mode", [&](const string& s) { config.setGaussMode(s); }, popsift::Config::getGaussModeUsage());
    val("desc-mode", [&](const string& s) { config.setDescMode(s); }, "Choice of descriptor extraction modes: loop, iloop, grid, igrid, notile");
    flag("popsift-mode", [&] { config.setMode(popsift::Config::PopSift); }, "PopSift extrema refinement (default)");
    flag("vlfeat-mode", [&] { config.setMode(popsift::Config::VLFeat); }, "VLFeat-like extrema refinement");
    flag("opencv-mode", [&] { config.setMode(popsift::Config::OpenCV); }, "OpenCV-like extrema refinement");
    flag("direct-scaling", [&] { config.setScalingMode(popsift::Config::ScaleDirect); }, "Direct each octave from upscaled orig instead of blurred level.");
    val("norm-multi", [&](const string& s) { config.setNormalizationMultiplier(atoi(s.c_str())); }, "Multiply the descriptor by pow(2,<int>).");
    val("norm-mode", [&](const string& s) { config.setNormMode(s); }, popsift::Config::getNormModeUsage());
    flag("root-sift", [&] { config.setNormMode(popsift::Config::RootSift); }, popsift::Config::getNormModeUsage());
    val("filter-max-extrema", [&](const string& s) { config.setFilterMaxExtrema(atoi(s.c_str())); }, "Approximate max number of extrema.");
    val("filter-grid", [&](const string& s) { config.setFilterGridSize(atoi(s.c_str())); }, "Grid edge length for extrema filtering");
    val("filter-sort", [&](const string& s) { config.setFilterSorting(s); }, "Sort extrema in each cell by scale: random, up or down");
    flag("print-gauss-tables", [&] { config.setPrintGaussTables(); }, "A debug output printing Gauss filter size and tables");
    flag("print-dev-info", [&] { print_dev_info = true; }, "A debug output printing CUDA device information");
    flag("print-time-info", [&] { print_time_info = true; }, "A debug output printing image processing time after load()");
    flag("write-as-uchar", [&] { write_as_uchar = true; }, "Output descriptors rounded to int.");
    flag("dont-write", [&] { dont_write = true; }, "Suppress descriptor output");
    flag("pgmread-loading", [&] { pgmread_loading = true; }, "Use the built-in PGM loader (always on here)");
    flag("float-mode", [&] { float_mode = true; }, "Upload image to GPU as float instead of byte");
    const map<string, string> shorts = {{"-h", "help"}, {"-v", "verbose"}, {"-l", "log"}, {"-i", "input-file"}};

    auto die = [&](const string& m) { cerr << "Error: " << m << endl << endl << "Usage:" << endl << endl; usage(o); exit(EXIT_FAILURE); };
    for (int i = 1; i < argc; i++) {
        string a = argv[i], name, value;
        bool has_value = false;
        if (a.rfind("--", 0) == 0) {
            name = a.substr(2);
            const size_t eq = name.find('=');
            if (eq != string::npos) { value = name.substr(eq + 1); name = name.substr(0, eq); has_value = true; }
        } else if (shorts.count(a)) name = shorts.at(a);
        else die("unrecognised option '" + a + "'");
        auto it = o.find(name);
        if (it == o.end()) die("unrecognised option '" + a + "'");
        if (it->second.has_arg && !has_value) {
            if (i + 1 >= argc) die("the required argument for option '--" + name + "' is missing");
            value = argv[++i];
        }
        it->second.fn(value);
    }
    if (inputFile.empty()) die("the option '--input-file' is required but missing");
}

static void collectFilenames(list<string>& inputFiles, const fs::path& dir)
{
    for (const auto& e : fs::directory_iterator(dir)) {
        if (e.is_regular_file()) inputFiles.push_back(e.path().string());
        else if (e.is_directory()) collectFilenames(inputFiles, e.path());
    }
}

static SiftJob* process_image(const string& inputFile, PopSift& sift)
{
    int w{}, h{};
    unsigned char* image_data = readPGMfile(inputFile, w, h);
    if (image_data == nullptr) exit(EXIT_FAILURE);
    cout << "Loading " << w << " x " << h << " image " << inputFile << endl;
    SiftJob* job;
    if (!float_mode) {
        job = sift.enqueue(w, h, image_data);
    } else {
        auto f = new float[(size_t)w * h];
        for (size_t i = 0; i < (size_t)w * h; i++) f[i] = float(image_data[i]) / 256.0f;   // reference main.cpp:234
        job = sift.enqueue(w, h, f);
        delete[] f;
    }
    delete[] image_data;
    return job;
}

static void read_job(SiftJob* job, bool really_write)
{
    popsift::Features* feature_list = job->get();
    cerr << "Number of feature points: " << feature_list->getFeatureCount()
         << " number of feature descriptors: " << feature_list->getDescriptorCount() << endl;
    if (really_write) {
        std::ofstream of("output-features.txt");
        feature_list->print(of, write_as_uchar);
    }
    delete feature_list;
}

int main(int argc, char** argv)
{
    popsift::Config config;
    list<string> inputFiles;
    string inputFile{};
    std::cout << "PopSift version: " << POPSIFT_VERSION_STRING << std::endl;
    try {
        parseargs(argc, argv, config, inputFile);
        std::cout << inputFile << std::endl;
    } catch (std::exception& e) {
        std::cout << e.what() << std::endl;
        return EXIT_FAILURE;
    }
    if (fs::exists(inputFile)) {
        if (fs::is_directory(inputFile)) {
            cout << inputFile << " is directory" << endl;
            collectFilenames(inputFiles, inputFile);
            if (inputFiles.empty()) { cerr << "No files in directory, nothing to do" << endl; return EXIT_SUCCESS; }
        } else if (fs::is_regular_file(inputFile)) {
            inputFiles.push_back(inputFile);
        } else {
            cout << "Input file is neither regular file nor directory, nothing to do" << endl;
            return EXIT_FAILURE;
        }
    }
    popsift::cuda::device_prop_t deviceInfo;
    deviceInfo.set(0, print_dev_info);
    if (print_dev_info) deviceInfo.print();

    PopSift sift(config, popsift::Config::ExtractingMode, float_mode ? PopSift::FloatImages : PopSift::ByteImages);
    std::queue<SiftJob*> jobs;
    for (const auto& f : inputFiles) jobs.push(process_image(f, sift));
    int rc = EXIT_SUCCESS;
    while (!jobs.empty()) {
        SiftJob* job = jobs.front();
        jobs.pop();
        if (job) {
            try { read_job(job, !dont_write); }
            catch (const std::exception& e) { cerr << "popsift-demo: " << e.what() << endl; rc = EXIT_FAILURE; }
            delete job;
        }
    }
    sift.uninit();
    return rc;
}
