"""Builds the product's native code IN-TREE with nvcc for sm_100a (no JIT cache).

    python -m popsift_b200.build            # libpopsift_b200.so + popsift-demo

Outputs (git-ignored, shipped to the GPU box by gpurun):
    popsift_b200/lib/libpopsift_b200.so     C-ABI (include/popsift_b200.h) + C++ API (include/popsift/*.h)
    popsift_b200/bin/popsift-demo           CLI mirroring the reference's popsift-demo
"""
from __future__ import annotations

import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "popsift_b200")
CSRC = os.path.join(PKG, "csrc")
LIB_DIR = os.path.join(PKG, "lib")
BIN_DIR = os.path.join(PKG, "bin")
OBJ_DIR = os.path.join(PKG, "build")
LIB = os.path.join(LIB_DIR, "libpopsift_b200.so")
DEMO = os.path.join(BIN_DIR, "popsift-demo")

ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
NVCC_FLAGS = ["-std=c++17", "-O3", "-lineinfo", "-Xcompiler", "-fPIC",
              "--fmad=false",   # no implicit contraction anywhere; every fma in the kernels is explicit
              "-I" + os.path.join(ROOT, "include"), "-I" + CSRC]

LIB_SOURCES = ["ps_tables.cpp", "ps_ctx.cu", "k_pyramid.cu", "k_pyramid_march.cu", "tma_util.cu", "k_match.cu", "k_extrema.cu", "k_filter.cu", "k_orient.cu", "k_desc.cu", "k_desc_modes.cu",
               "host/sift_conf.cpp", "host/features.cpp", "host/popsift.cpp", "host/device_prop.cpp", "host/log_dump.cpp"]
DEMO_SOURCES = ["app/popsift_demo.cpp", "app/pgmread.cpp"]
MATCH_SOURCES = ["app/popsift_match.cpp"]            # + app/pgmread.cpp
MATCH = os.path.join(BIN_DIR, "popsift-match")
API_CHECK = os.path.join(BIN_DIR, "api_check")
API_CHECK_SRC = os.path.join(ROOT, "tests", "cpp", "api_check.cpp")
API_BENCH = os.path.join(BIN_DIR, "api_bench")
API_BENCH_SRC = os.path.join(ROOT, "tests", "cpp", "api_bench.cpp")
PART_CHECK = os.path.join(BIN_DIR, "partition_check")
PART_CHECK_SRC = os.path.join(ROOT, "tests", "cpp", "partition_check.cpp")
PGM_CHECK = os.path.join(BIN_DIR, "pgmread_check")
PGM_CHECK_SRC = os.path.join(ROOT, "tests", "cpp", "pgmread_check.cpp")


def _newer(target: str, deps) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps if os.path.exists(d))


def _headers():
    out = []
    for base in (os.path.join(ROOT, "include"), CSRC):
        for dp, _, fs in os.walk(base):
            out += [os.path.join(dp, f) for f in fs if f.endswith((".h", ".hpp", ".cuh"))]
    return out


def build(verbose: bool = False, force: bool = False) -> str:
    for d in (LIB_DIR, BIN_DIR, OBJ_DIR):
        os.makedirs(d, exist_ok=True)
    hdrs = _headers()
    objs = []
    procs = []
    for src in LIB_SOURCES + DEMO_SOURCES + MATCH_SOURCES:
        sp = os.path.join(CSRC, src)
        if not os.path.exists(sp):
            continue
        obj = os.path.join(OBJ_DIR, src.replace("/", "_") + ".o")
        objs.append((src, obj))
        if force or _newer(obj, [sp] + hdrs):
            cmd = ["nvcc"] + ARCH + NVCC_FLAGS + ["-x", "cu", "-c", sp, "-o", obj]
            if verbose:
                cmd.insert(1, "-Xptxas=-v")
                print(" ".join(cmd))
            procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write(out.decode(errors="replace"))
        if p.returncode != 0:
            raise RuntimeError("nvcc failed for %s" % src)
    lib_objs = [o for s, o in objs if s in LIB_SOURCES]
    demo_objs = [o for s, o in objs if s in DEMO_SOURCES]
    if force or _newer(LIB, lib_objs):
        subprocess.check_call(["nvcc"] + ARCH + ["-shared", "-o", LIB] + lib_objs + ["-lpthread"])
    if demo_objs and (force or _newer(DEMO, demo_objs + [LIB])):
        subprocess.check_call(["nvcc"] + ARCH + ["-o", DEMO] + demo_objs +
                              ["-L" + LIB_DIR, "-lpopsift_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../lib", "-lpthread"])
    match_objs = [o for s, o in objs if s in MATCH_SOURCES or s == "app/pgmread.cpp"]
    if len(match_objs) == 2 and (force or _newer(MATCH, match_objs + [LIB])):
        subprocess.check_call(["nvcc"] + ARCH + ["-o", MATCH] + match_objs +
                              ["-L" + LIB_DIR, "-lpopsift_b200", "-Xlinker", "-rpath", "-Xlinker", "$ORIGIN/../lib", "-lpthread"])
    if os.path.exists(API_CHECK_SRC) and (force or _newer(API_CHECK, [API_CHECK_SRC, LIB] + hdrs)):
        # a plain host compiler is enough for a caller of the C++ API: no CUDA in the translation unit
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), API_CHECK_SRC, "-o", API_CHECK,
                               "-L" + LIB_DIR, "-lpopsift_b200", "-Wl,-rpath," + "$ORIGIN/../lib", "-lpthread"])
    if os.path.exists(API_BENCH_SRC) and (force or _newer(API_BENCH, [API_BENCH_SRC, LIB] + hdrs)):
        # end-to-end timing through the C++ API from pageable frames (bench.py's e2e leg)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(ROOT, "include"), API_BENCH_SRC, "-o", API_BENCH,
                               "-L" + LIB_DIR, "-lpopsift_b200", "-Wl,-rpath," + "$ORIGIN/../lib", "-lpthread"])
    part_hdr = os.path.join(CSRC, "k_partition.h")
    if os.path.exists(PART_CHECK_SRC) and (force or _newer(PART_CHECK, [PART_CHECK_SRC, part_hdr])):
        # host-only sweep of the marching kernels' work partition (no CUDA, no library)
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + CSRC, PART_CHECK_SRC, "-o", PART_CHECK])
    pgm_src = os.path.join(CSRC, "app", "pgmread.cpp")
    if os.path.exists(PGM_CHECK_SRC) and (force or _newer(PGM_CHECK, [PGM_CHECK_SRC, pgm_src, os.path.join(CSRC, "app", "pgmread.h")])):
        # host-only check of popsift-demo's image reader
        subprocess.check_call(["g++", "-std=c++17", "-O2", "-I" + os.path.join(CSRC, "app"), PGM_CHECK_SRC, pgm_src, "-o", PGM_CHECK])
    return LIB


if __name__ == "__main__":
    build(verbose="-v" in sys.argv, force="-f" in sys.argv)
    print(LIB)
