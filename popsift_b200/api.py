"""Python host-side mirror of the reference's public interface, over the C ABI.

The product is libpopsift_b200.so (C ABI in include/popsift_b200.h, C++ drop-in API in
include/popsift/*.h).  This module binds the C ABI with ctypes and offers the same vocabulary as
the reference (popsift::Config setters, PopSift.enqueue -> SiftJob.get -> Features) so that the
parity tests read like the reference's own usage (reference src/application/main.cpp:172-264).

There is no CPU fallback here: if the shared library is missing, or no CUDA device is usable,
construction raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("POPSIFT_B200_LIB") or os.path.join(_HERE, "lib", "libpopsift_b200.so")

PS_MAX_OCTAVES = 20
MODE = {"popsift": 0, "opencv": 1, "vlfeat": 2}
NORM = {"rootsift": 0, "RootSift": 0, "classic": 1}
GAUSS = {"vlfeat": 0, "vlfeat-hw-interpolated": 1, "relative": 1, "vlfeat-direct": 2, "opencv": 3, "fixed9": 4, "fixed15": 5}
DESC = {"loop": 0, "iloop": 1, "grid": 2, "igrid": 3, "notile": 4}
FILTER_SORT = {"random": 0, "down": 1, "up": 2}
STAGES = ("h2d", "pyramid", "extrema", "orientation", "descriptors", "total")


class PsConfig(C.Structure):
    _fields_ = [("octaves", C.c_int32), ("levels", C.c_int32), ("sigma", C.c_float), ("edge_limit", C.c_float),
                ("threshold", C.c_float), ("upscale", C.c_float), ("initial_blur", C.c_float),
                ("has_initial_blur", C.c_int32), ("sift_mode", C.c_int32), ("gauss_mode", C.c_int32),
                ("desc_mode", C.c_int32), ("norm_mode", C.c_int32), ("norm_multi", C.c_int32),
                ("max_extrema", C.c_int32), ("scaling_mode", C.c_int32), ("filter_max_extrema", C.c_int32),
                ("filter_grid_size", C.c_int32), ("filter_sort", C.c_int32)]


class PsGaussTables(C.Structure):
    _fields_ = [("inc_filter", C.c_float * (12 * 32)), ("inc_sigma", C.c_float * 12), ("inc_span", C.c_int32 * 12),
                ("dd_filter0", C.c_float * 32), ("dd_sigma0", C.c_float), ("dd_span0", C.c_int32),
                ("peak_threshold", C.c_float), ("sigma_k", C.c_float),
                ("dd_filter", C.c_float * (20 * 32)), ("dd_sigma", C.c_float * 20), ("dd_span", C.c_int32 * 20),
                ("abs_filter", C.c_float * (12 * 32)), ("abs_sigma", C.c_float * 12), ("abs_span", C.c_int32 * 12),
                ("inc_ifilter", C.c_float * (12 * 32)), ("inc_ispan", C.c_int32 * 12),
                ("absn_filter", C.c_float * (12 * 32)), ("absn_sigma", C.c_float * 12), ("absn_span", C.c_int32 * 12)]


FEATURE_DTYPE = np.dtype([("octave", "<i4"), ("x", "<f4"), ("y", "<f4"), ("sigma", "<f4"), ("num_ori", "<i4"),
                          ("ori", "<f4", (4,)), ("pad", "<i4"), ("desc_ptr", "<u8", (4,))])
EXTREMUM_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("lpos", "<i4"), ("sigma", "<f4"), ("octave", "<i4"),
                           ("num_ori", "<i4"), ("idx_ori", "<i4"), ("ori", "<f4", (4,))])
assert FEATURE_DTYPE.itemsize == 72 and EXTREMUM_DTYPE.itemsize == 44

EXPORTS = ["ps_abi_version", "ps_config_default", "ps_gauss_tables_compute", "ps_geometry", "ps_create", "ps_destroy",
           "ps_last_error", "ps_submit_u8", "ps_submit_f32", "ps_submit_dev_u8", "ps_counts", "ps_download",
           "ps_sync", "ps_debug_plane", "ps_debug_extrema", "ps_slot_geometry", "ps_set_timing", "ps_stage_ms",
           "ps_launch_count", "ps_slot_stream", "ps_run_pyramid_only", "ps_run_level_only", "ps_host_alloc", "ps_host_free",
           "ps_download_dev", "ps_dev_alloc", "ps_dev_free", "ps_dev_to_host", "ps_wait_input", "ps_match", "ps_pointer_device", "ps_host_to_dev",
           "ps_debug_level0_plan", "ps_format_gauss_tables"]

_lib = None


def load_library():
    """dlopen the product library; raises if it has not been built (python -m popsift_b200.build)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError("popsift_b200: %s is missing - build it with `python -m popsift_b200.build` "
                           "(there is no fallback path)" % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    L.ps_abi_version.restype = C.c_int
    L.ps_config_default.argtypes = [C.POINTER(PsConfig)]
    L.ps_gauss_tables_compute.argtypes = [C.POINTER(PsConfig), C.POINTER(PsGaussTables)]
    L.ps_geometry.argtypes = [C.POINTER(PsConfig), C.c_int, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.ps_create.restype = C.c_void_p
    L.ps_create.argtypes = [C.c_int, C.POINTER(PsConfig), C.c_int, C.c_int, C.c_int]
    L.ps_destroy.argtypes = [C.c_void_p]
    L.ps_last_error.restype = C.c_char_p
    L.ps_last_error.argtypes = [C.c_void_p]
    L.ps_submit_u8.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.ps_submit_f32.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int]
    L.ps_submit_dev_u8.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_int, C.c_int]
    L.ps_counts.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.ps_download.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.ps_sync.argtypes = [C.c_void_p, C.c_int]
    L.ps_wait_input.argtypes = [C.c_void_p, C.c_int]
    L.ps_match.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.ps_pointer_device.argtypes = [C.c_void_p]
    L.ps_host_to_dev.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    L.ps_debug_plane.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
    L.ps_debug_extrema.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
    L.ps_debug_level0_plan.argtypes = [C.c_int, C.c_int, C.c_int, C.c_int, C.c_float, C.c_int]
    L.ps_format_gauss_tables.argtypes = [C.c_void_p, C.c_char_p, C.c_size_t]
    L.ps_slot_geometry.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    L.ps_set_timing.argtypes = [C.c_void_p, C.c_int]
    L.ps_stage_ms.argtypes = [C.c_void_p, C.c_int, C.POINTER(C.c_float)]
    L.ps_launch_count.restype = C.c_int64
    L.ps_launch_count.argtypes = [C.c_void_p]
    L.ps_slot_stream.restype = C.c_void_p
    L.ps_slot_stream.argtypes = [C.c_void_p, C.c_int]
    L.ps_run_pyramid_only.argtypes = [C.c_void_p, C.c_int]
    L.ps_run_level_only.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
    L.ps_host_alloc.restype = C.c_void_p
    L.ps_host_alloc.argtypes = [C.c_size_t]
    L.ps_host_free.argtypes = [C.c_void_p]
    L.ps_download_dev.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.ps_dev_alloc.restype = C.c_void_p
    L.ps_dev_alloc.argtypes = [C.c_size_t]
    L.ps_dev_free.argtypes = [C.c_void_p]
    L.ps_dev_to_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    _lib = L
    return L


class PopSiftError(RuntimeError):
    pass


class Config:
    """popsift::Config (reference src/popsift/sift_conf.h:29-409): same setters, same defaults."""

    # enums, reference sift_conf.h:33-107
    PopSift, OpenCV, VLFeat = 0, 1, 2
    RootSift, Classic = 0, 1
    ExtractingMode, MatchingMode = 0, 1

    def __init__(self):
        self._c = PsConfig()
        load_library().ps_config_default(C.byref(self._c))
        self.verbose = False
        self._log_mode = 0
        self._print_gauss_tables = False

    # public fields of the reference struct
    octaves = property(lambda s: s._c.octaves, lambda s, v: setattr(s._c, "octaves", int(v)))
    levels = property(lambda s: s._c.levels, lambda s, v: setattr(s._c, "levels", int(v)))
    sigma = property(lambda s: s._c.sigma, lambda s, v: setattr(s._c, "sigma", float(v)))
    _edge_limit = property(lambda s: s._c.edge_limit, lambda s, v: setattr(s._c, "edge_limit", float(v)))

    def setMode(self, m):
        self._c.sift_mode = MODE[m] if isinstance(m, str) else int(m)

    def setGaussMode(self, m):
        if isinstance(m, str):
            if m not in GAUSS:
                raise PopSiftError("Bad Gauss mode.")          # reference sift_conf.cu:101
            m = GAUSS[m]
        self._c.gauss_mode = int(m)

    def setDescMode(self, m="loop"):
        if isinstance(m, str):
            if m not in DESC:
                raise PopSiftError("specified descriptor extraction mode must be one of loop, grid or igrid")
            m = DESC[m]
        self._c.desc_mode = int(m)

    def setNormMode(self, m):
        if isinstance(m, str):
            if m not in ("RootSift", "classic"):
                raise PopSiftError("Bad Normalization mode.")   # reference sift_conf.cu:202
            m = NORM[m]
        self._c.norm_mode = int(m)

    def setUseRootSift(self, on):
        self._c.norm_mode = 0 if on else 1

    def getUseRootSift(self):
        return self._c.norm_mode == 0

    def setNormalizationMultiplier(self, m):
        self._c.norm_multi = int(m)

    def getNormalizationMultiplier(self):
        return self._c.norm_multi

    def setDownsampling(self, v):
        self._c.upscale = -float(v)

    def getUpscaleFactor(self):
        return self._c.upscale

    def setOctaves(self, v): self._c.octaves = int(v)
    def setLevels(self, v): self._c.levels = int(v)
    def setSigma(self, v): self._c.sigma = float(v)
    def setEdgeLimit(self, v): self._c.edge_limit = float(v)
    def setThreshold(self, v): self._c.threshold = float(v)
    def setVerbose(self, on=True): self.verbose = bool(on)
    def setLogMode(self, mode=1): self._log_mode = int(mode)
    def setFilterMaxExtrema(self, n): self._c.filter_max_extrema = int(n)
    def setFilterGridSize(self, n): self._c.filter_grid_size = int(n)
    def getFilterMaxExtrema(self): return self._c.filter_max_extrema
    def getFilterGridSize(self): return self._c.filter_grid_size

    def setFilterSorting(self, m):
        if isinstance(m, str):
            if m not in FILTER_SORT:
                raise PopSiftError("filter sorting mode must be one of up, down or random")   # reference sift_conf.cu:141
            m = FILTER_SORT[m]
        self._c.filter_sort = int(m)

    def setScalingMode(self, m=1):
        """Config::ScaleDirect = 0, Config::ScaleDefault = 1 (reference sift_conf.h:75-80)"""
        self._c.scaling_mode = int(m)
    def setPrintGaussTables(self): self._print_gauss_tables = True

    def gauss_tables_text(self) -> str:
        """what the reference prints under setPrintGaussTables() (gauss_filter.cu:24-121,146-161)"""
        L = load_library()
        n = L.ps_format_gauss_tables(C.byref(self._c), None, 0)
        if n < 0:
            raise PopSiftError("unsupported configuration (sigma > 2.0 or levels > 12)")
        buf = C.create_string_buffer(n + 1)
        L.ps_format_gauss_tables(C.byref(self._c), buf, n + 1)
        return buf.value.decode()

    def setInitialBlur(self, blur):
        self._c.has_initial_blur = 0 if blur == 0.0 else 1      # reference sift_conf.cu:246-255
        self._c.initial_blur = float(blur)

    def hasInitialBlur(self): return bool(self._c.has_initial_blur)
    def getInitialBlur(self): return self._c.initial_blur
    def getSiftMode(self): return self._c.sift_mode
    def getMaxExtrema(self): return self._c.max_extrema

    def getPeakThreshold(self):
        return float(np.float32(self._c.threshold) * np.float32(0.5) * np.float32(255.0) / np.float32(self._c.levels))

    def equal(self, other: "Config") -> bool:
        """reference sift_conf.cu:286-304 (desc mode, filter settings, log mode, verbose not compared)"""
        a, b = self._c, other._c
        return all(getattr(a, f) == getattr(b, f) for f in
                   ("octaves", "levels", "sigma", "edge_limit", "threshold", "upscale", "max_extrema", "gauss_mode",
                    "sift_mode", "has_initial_blur", "initial_blur", "norm_mode", "norm_multi"))

    __eq__ = equal

    def gauss_tables(self) -> PsGaussTables:
        t = PsGaussTables()
        rc = load_library().ps_gauss_tables_compute(C.byref(self._c), C.byref(t))
        if rc != 0:
            raise PopSiftError("unsupported configuration (sigma > 2.0, levels > 12, gauss mode)")
        return t

    def geometry(self, w: int, h: int):
        W = (C.c_int32 * PS_MAX_OCTAVES)()
        H = (C.c_int32 * PS_MAX_OCTAVES)()
        n = load_library().ps_geometry(C.byref(self._c), w, h, W, H)
        if n < 1:
            raise PopSiftError("bad geometry")
        return [(W[i], H[i]) for i in range(n)]


class _PinnedBlock:
    """One page-locked result buffer (features + descriptors); returned to its pool when released."""

    def __init__(self, lib, n_feat: int, n_desc: int):
        self.lib, self.n_feat, self.n_desc = lib, n_feat, n_desc
        self.pf = lib.ps_host_alloc(max(n_feat, 1) * 72)
        self.pd = lib.ps_host_alloc(max(n_desc, 1) * 512)
        if not self.pf or not self.pd:
            raise PopSiftError("ps_host_alloc failed")
        self.feat = np.ctypeslib.as_array(C.cast(self.pf, C.POINTER(C.c_uint8)), shape=(max(n_feat, 1) * 72,)).view(FEATURE_DTYPE)
        self.desc = np.ctypeslib.as_array(C.cast(self.pd, C.POINTER(C.c_float)), shape=(max(n_desc, 1), 128))

    def free(self):
        if self.pf:
            self.lib.ps_host_free(self.pf); self.lib.ps_host_free(self.pd)
            self.pf = self.pd = None


class _Pool:
    """Page-locked result buffers of one PopSift, recycled between images."""

    def __init__(self):
        self.blocks, self.closed = [], False

    def close(self):
        self.closed = True
        while self.blocks:
            self.blocks.pop().free()


class _Lease:
    """Ties a page-locked block to the numpy arrays that view it: every array handed out has a ctypes
    buffer as its base, the buffers hold this object, and only when the LAST array is gone does the block
    go back to its pool (or is freed, if the PopSift it came from has been shut down meanwhile).  So
    `d = ps.enqueue(...).get().desc` stays valid for as long as `d` lives."""

    def __init__(self, pool: _Pool, block: _PinnedBlock):
        self._pool, self._block = pool, block

    def __del__(self):
        blk, self._block = self._block, None
        if blk is None:
            return
        if self._pool.closed:
            blk.free()
        else:
            self._pool.blocks.append(blk)


def _leased_views(pool: _Pool, blk: _PinnedBlock, nf: int, nd: int):
    lease = _Lease(pool, blk)
    fbuf = (C.c_uint8 * (nf * 72)).from_address(blk.pf)
    dbuf = (C.c_float * (nd * 128)).from_address(blk.pd)
    fbuf._lease = dbuf._lease = lease
    feat = np.frombuffer(fbuf, dtype=FEATURE_DTYPE, count=nf) if nf else np.zeros(0, FEATURE_DTYPE)
    desc = np.frombuffer(dbuf, dtype=np.float32, count=nd * 128).reshape(nd, 128) if nd else np.zeros((0, 128), np.float32)
    return feat, desc


class Features:
    """popsift::FeaturesHost (reference src/popsift/features.h:71-102).  `feat` / `desc` view a page-locked
    buffer that stays alive as long as any array that views it (see _Lease)."""

    def __init__(self, feat: np.ndarray, desc: np.ndarray):
        self.feat = feat
        self.desc = desc
        base = desc.ctypes.data if len(desc) else 0
        idx = np.full((len(feat), 4), -1, dtype=np.int64)
        for k in range(4):
            m = feat["num_ori"] > k
            idx[m, k] = (feat["desc_ptr"][m, k].astype(np.int64) - base) // 512
        self.desc_idx = idx

    def getFeatureCount(self): return len(self.feat)
    def getDescriptorCount(self): return len(self.desc)
    size = getFeatureCount

    def keypoints(self):
        """rows (x, y, sigma, theta) per (feature, orientation) and the matching descriptors"""
        rows, idx = [], []
        for i, f in enumerate(self.feat):
            for k in range(int(f["num_ori"])):
                rows.append((f["x"], f["y"], f["sigma"], f["ori"][k]))
                idx.append(int(self.desc_idx[i, k]))
        kp = np.array(rows, dtype=np.float64).reshape(-1, 4)
        return kp, (self.desc[np.array(idx, dtype=np.int64)] if idx else self.desc[:0])

    def print(self, fh, write_as_uchar=False):
        """reference Feature::print (features.cu:310-330): x y 1/s^2 0 1/s^2 d0..d127"""
        for i, f in enumerate(self.feat):
            sigval = np.float32(1.0) / (np.float32(f["sigma"]) * np.float32(f["sigma"]))
            for k in range(int(f["num_ori"])):
                d = self.desc[self.desc_idx[i, k]]
                if write_as_uchar:
                    ds = " ".join("%g" % np.round(v) for v in d)
                else:
                    ds = " ".join("%.3g" % v for v in d)
                fh.write("%g %g %g 0 %g %s \n" % (f["x"], f["y"], sigval, sigval, ds))


class FeaturesDev:
    """FeaturesDev (reference src/popsift/features.h:104-122): Feature records, descriptors and the
    descriptor -> feature reverse map in DEVICE memory owned by this object (Config.MatchingMode).
    getFeatures / getDescriptors / getReverseMap return raw device addresses."""

    def __init__(self, lib, n_feat: int, n_desc: int):
        self._lib, self._nf, self._nd = lib, n_feat, n_desc
        self._pf = lib.ps_dev_alloc(max(n_feat, 1) * 72)
        self._pd = lib.ps_dev_alloc(max(n_desc, 1) * 512)
        self._pr = lib.ps_dev_alloc(max(n_desc, 1) * 4)
        if not (self._pf and self._pd and self._pr):
            raise PopSiftError("Failed to allocate device memory for features")

    def getFeatureCount(self): return self._nf
    def getDescriptorCount(self): return self._nd
    size = getFeatureCount
    def getFeatures(self): return self._pf
    def getDescriptors(self): return self._pd
    def getReverseMap(self): return self._pr

    MATCH_AUTO, MATCH_EXACT, MATCH_TENSOR = 0, 1, 2

    def match(self, other: "FeaturesDev", flags: int = 0) -> np.ndarray:
        """FeaturesDev::match (reference features.cu:282-304) without the printing: (n_desc, 3) int32 rows
        (best index, second index, accept) of every descriptor of self among other's (C ABI ps_match)."""
        return match_descriptors_dev(self._lib, self._pd, self._nd, other._pd, other._nd, flags)

    def to_host(self):
        """(features, descriptors, reverse map) copied to numpy arrays; desc_ptr still holds DEVICE addresses"""
        feat = np.zeros(self._nf, FEATURE_DTYPE)
        desc = np.zeros((self._nd, 128), np.float32)
        rev = np.zeros(self._nd, np.int32)
        for arr, ptr in ((feat, self._pf), (desc, self._pd), (rev, self._pr)):
            if arr.nbytes and self._lib.ps_dev_to_host(arr.ctypes.data, ptr, arr.nbytes) != 0:
                raise PopSiftError("device -> host copy failed")
        return feat, desc, rev

    def __del__(self):
        lib = getattr(self, "_lib", None)
        if lib is not None:
            for p in (self._pf, self._pd, self._pr):
                if p:
                    lib.ps_dev_free(p)
            self._pf = self._pd = self._pr = None


def match_descriptors_dev(lib, d_left, n_left, d_right, n_right, flags=0, device=None) -> np.ndarray:
    out = np.zeros((max(n_left, 0), 3), np.int32)
    if n_left <= 0:
        return out
    d_out = lib.ps_dev_alloc(out.nbytes)
    if not d_out:
        raise PopSiftError("ps_dev_alloc failed")
    try:
        dev = device if device is not None else max(0, lib.ps_pointer_device(d_left))
        rc = lib.ps_match(dev, d_left, n_left, d_right, n_right, d_out, flags)
        if rc != 0:
            raise PopSiftError("ps_match error %d: %s" % (rc, lib.ps_last_error(None).decode()))
        if lib.ps_dev_to_host(out.ctypes.data, d_out, out.nbytes) != 0:
            raise PopSiftError("device -> host copy failed")
    finally:
        lib.ps_dev_free(d_out)
    return out


def match_descriptors(left: np.ndarray, right: np.ndarray, flags=0, device=0) -> np.ndarray:
    """host arrays (n, 128) float32 -> (n_left, 3) int32 (best, second, accept); copies through device memory"""
    lib = load_library()
    left = np.ascontiguousarray(left, np.float32); right = np.ascontiguousarray(right, np.float32)
    import ctypes as _C
    bufs = []
    try:
        for a in (left, right):
            p = lib.ps_dev_alloc(max(a.nbytes, 4))
            if not p:
                raise PopSiftError("ps_dev_alloc failed")
            bufs.append(p)
            if a.nbytes and lib.ps_host_to_dev(p, a.ctypes.data, a.nbytes) != 0:
                raise PopSiftError("host -> device copy failed")
        return match_descriptors_dev(lib, bufs[0], len(left), bufs[1], len(right), flags, device)
    finally:
        for p in bufs:
            lib.ps_dev_free(p)


class SiftJob:
    """SiftJob (reference src/popsift/popsift.h:44-100): get() blocks until the features are there."""

    def __init__(self, owner: "PopSift", slot: int):
        self._owner, self._slot, self._result = owner, slot, None

    def getBase(self):
        if self._result is None:
            self._result = self._owner._collect(self._slot)
        return self._result

    def get(self) -> Features:
        """host features (None under Config.MatchingMode, like the reference's dynamic_cast)"""
        r = self.getBase()
        return r if isinstance(r, Features) else None

    getHost = get

    def getDev(self) -> "FeaturesDev":
        r = self.getBase()
        return r if isinstance(r, FeaturesDev) else None


class PopSift:
    """PopSift (reference src/popsift/popsift.h:105-317) over the C ABI: one context per device,
    `slots` images in flight (the reference has one)."""

    ByteImages, FloatImages = 0, 1

    def __init__(self, config: Optional[Config] = None, mode=Config.ExtractingMode, imode=0, device: int = 0,
                 max_width: int = 0, max_height: int = 0, slots: int = 2):
        self._lib = load_library()
        self._config = config or Config()
        self._imode, self._device, self._nslots = imode, device, slots
        self._mode = mode
        self._ctx = None
        self._max = (max_width, max_height)
        self._next = 0
        self._busy = [None] * slots
        self._keep = [None] * slots      # the image each slot's DMA may still be reading (page-locked inputs)
        self._pool = _Pool()     # page-locked result buffers, recycled when the arrays that view them die
        if max_width and max_height:
            self._create(max_width, max_height)

    def _create(self, w, h):
        ctx = self._lib.ps_create(self._device, C.byref(self._config._c), w, h, self._nslots)
        if not ctx:
            raise PopSiftError(self._lib.ps_last_error(None).decode())
        self._ctx = ctx
        self._max = (w, h)

    def _check(self, rc):
        if rc != 0:
            raise PopSiftError("popsift_b200 error %d: %s" % (rc, self._lib.ps_last_error(self._ctx).decode()))

    def enqueue(self, w: int, h: int, image) -> SiftJob:
        img = np.ascontiguousarray(image)
        want = np.uint8 if self._imode == self.ByteImages else np.float32
        if img.dtype != want:
            raise PopSiftError("Image mode error")    # reference popsift.cpp:247-253
        assert img.size == w * h
        if self._ctx is None:
            self._create(w, h)
        if w > self._max[0] or h > self._max[1]:
            raise PopSiftError("image larger than the context was created for")
        slot = self._next
        self._next = (self._next + 1) % self._nslots
        if self._busy[slot] is not None:
            self._busy[slot].getBase()   # FIFO: the slot's previous job completes first
        fn = self._lib.ps_submit_u8 if self._imode == self.ByteImages else self._lib.ps_submit_f32
        self._keep[slot] = img           # released when the slot's next image is submitted (its job has completed by then)
        self._check(fn(self._ctx, slot, img.ctypes.data, w, h))
        job = SiftJob(self, slot)
        self._busy[slot] = job
        return job

    def _collect(self, slot: int) -> Features:
        nf, nd = C.c_int32(), C.c_int32()
        self._check(self._lib.ps_counts(self._ctx, slot, C.byref(nf), C.byref(nd)))
        if self._mode == Config.MatchingMode:
            fd = FeaturesDev(self._lib, nf.value, nd.value)
            self._check(self._lib.ps_download_dev(self._ctx, slot, fd.getFeatures(), fd.getDescriptors(), fd.getReverseMap()))
            self._busy[slot] = None
            return fd
        blk = None
        for i, b in enumerate(self._pool.blocks):
            if b.n_feat >= nf.value and b.n_desc >= nd.value:
                blk = self._pool.blocks.pop(i)
                break
        if blk is None:
            blk = _PinnedBlock(self._lib, int(nf.value * 1.25) + 1024, int(nd.value * 1.25) + 1024)
        rc = self._lib.ps_download(self._ctx, slot, blk.pf, blk.pd)
        feat, desc = _leased_views(self._pool, blk, nf.value, nd.value)     # from here on the lease owns the block
        self._check(rc)
        self._busy[slot] = None
        return Features(feat, desc)

    # --- test / benchmark taps ------------------------------------------------------------
    def plane(self, slot, octave, level, dog=False) -> np.ndarray:
        n = C.c_int32()
        W = (C.c_int32 * PS_MAX_OCTAVES)()
        H = (C.c_int32 * PS_MAX_OCTAVES)()
        self._check(self._lib.ps_slot_geometry(self._ctx, slot, C.byref(n), W, H))
        out = np.zeros((H[octave], W[octave]), dtype=np.float32)
        self._check(self._lib.ps_debug_plane(self._ctx, slot, octave, level, 1 if dog else 0, out.ctypes.data))
        return out

    def slot_geometry(self, slot):
        n = C.c_int32()
        W = (C.c_int32 * PS_MAX_OCTAVES)()
        H = (C.c_int32 * PS_MAX_OCTAVES)()
        self._check(self._lib.ps_slot_geometry(self._ctx, slot, C.byref(n), W, H))
        return [(W[i], H[i]) for i in range(n.value)]

    def extrema(self, slot) -> np.ndarray:
        n = self._lib.ps_debug_extrema(self._ctx, slot, None, 0)
        out = np.zeros(max(n, 0), dtype=EXTREMUM_DTYPE)
        if n > 0:
            self._lib.ps_debug_extrema(self._ctx, slot, out.ctypes.data, n)
        return out

    def set_timing(self, on=True):
        self._check(self._lib.ps_set_timing(self._ctx, 1 if on else 0))

    def stage_ms(self, slot):
        ms = (C.c_float * 6)()
        self._check(self._lib.ps_stage_ms(self._ctx, slot, ms))
        return dict(zip(STAGES, [float(v) for v in ms]))

    def launch_count(self) -> int:
        return int(self._lib.ps_launch_count(self._ctx))

    def uninit(self):
        if self._ctx:
            self._lib.ps_destroy(self._ctx)
            self._ctx = None
        self._pool.close()        # frees the idle blocks; blocks still viewed by arrays are freed when those die

    def __del__(self):
        try:
            self.uninit()
        except Exception:
            pass
