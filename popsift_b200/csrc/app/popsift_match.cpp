// popsift-match for the Blackwell-native drop-in: same command line and console output as the reference's matcher
// application (reference src/application/match.cpp:41-279): extract both images under Config::MatchingMode (results
// stay on the device), then FeaturesDev::match prints one accept/reject line per left descriptor.
#include <popsift/common/device_prop.h>
#include <popsift/features.h>
#include <popsift/popsift.h>
#include <popsift/sift_conf.h>
#include <popsift/sift_config.h>
#include <popsift/version.hpp>

#include "pgmread.h"

#include <cstdlib>
#include <filesystem>
#include <functional>
#include <iostream>
#include <map>
#include <string>

using namespace std;
namespace fs = std::filesystem;

static bool print_dev_info = false;

struct Opt { bool has_arg; function<void(const string&)> fn; string help; };

static void usage(const map<string, Opt>& opts)
{
    cout << "Allowed options:" << endl;
    for (auto& kv : opts) cout << "  --" << kv.first << (kv.second.has_arg ? " arg" : "") << "\t" << kv.second.help << endl;
}

static void parseargs(int argc, char** argv, popsift::Config& config, string& lFile, string& rFile)
{
    map<string, Opt> o;
    auto flag = [&](const string& n, function<void()> f, const string& h) { o[n] = {false, [f](const string&) { f(); }, h}; };
    auto val = [&](const string& n, function<void(const string&)> f, const string& h) { o[n] = {true, f, h}; };
    flag("help", [&] { usage(o); exit(1); }, "Print usage");
    flag("verbose", [&] { config.setVerbose(); }, "");
    flag("log", [&] { config.setLogMode(popsift::Config::All); }, "Write debugging files");
    val("left", [&](const string& s) { lFile = s; }, "\"Left\"  input file");
    val("right", [&](const string& s) { rFile = s; }, "\"Right\" input file");
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
    flag("print-time-info", [] {}, "A debug output printing image processing time after load()");
    flag("write-as-uchar", [] {}, "Output descriptors rounded to int.");
    flag("dont-write", [] {}, "Suppress descriptor output");
    flag("pgmread-loading", [] {}, "Use the built-in PGM loader (always on here)");
    const map<string, string> shorts = {{"-h", "help"}, {"-v", "verbose"}, {"-l", "left"}, {"-r", "right"}};

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
    if (lFile.empty()) die("the option '--left' is required but missing");
    if (rFile.empty()) die("the option '--right' is required but missing");
}

static SiftJob* process_image(const string& inputFile, PopSift& sift)
{
    int w{}, h{};
    unsigned char* image_data = readPGMfile(inputFile, w, h);
    if (image_data == nullptr) exit(EXIT_FAILURE);
    SiftJob* job = sift.enqueue(w, h, image_data);
    delete[] image_data;
    return job;
}

int main(int argc, char** argv)
{
    popsift::Config config;
    string lFile{}, rFile{};
    std::cout << "PopSift version: " << POPSIFT_VERSION_STRING << std::endl;
    try {
        parseargs(argc, argv, config, lFile, rFile);
        std::cout << lFile << " <-> " << rFile << std::endl;
    } catch (std::exception& e) {
        std::cout << e.what() << std::endl;
        return EXIT_SUCCESS;
    }
    for (const string& f : {lFile, rFile})
        if (fs::exists(f) && !fs::is_regular_file(f)) {
            cout << "Input file " << f << " is not a regular file, nothing to do" << endl;
            return EXIT_FAILURE;
        }
    popsift::cuda::device_prop_t deviceInfo;
    deviceInfo.set(0, print_dev_info);
    if (print_dev_info) deviceInfo.print();

    int rc = EXIT_SUCCESS;
    PopSift sift(config, popsift::Config::MatchingMode);
    SiftJob* lJob = process_image(lFile, sift);
    SiftJob* rJob = process_image(rFile, sift);
    try {
        popsift::FeaturesDev* lFeatures = lJob->getDev();
        cout << "Number of features:    " << lFeatures->getFeatureCount() << endl;
        cout << "Number of descriptors: " << lFeatures->getDescriptorCount() << endl;
        popsift::FeaturesDev* rFeatures = rJob->getDev();
        cout << "Number of features:    " << rFeatures->getFeatureCount() << endl;
        cout << "Number of descriptors: " << rFeatures->getDescriptorCount() << endl;
        cout.flush();
        lFeatures->match(rFeatures);
        delete lFeatures;
        delete rFeatures;
    } catch (const std::exception& e) {
        cerr << "popsift-match: " << e.what() << endl;
        rc = EXIT_FAILURE;
    }
    delete lJob;
    delete rJob;
    sift.uninit();
    return rc;
}
