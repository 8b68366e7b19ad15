"""Small workloads through every kernel family, for compute-sanitizer (memcheck / racecheck):
    compute-sanitizer --tool memcheck python tools/sanitize_run.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from popsift_b200 import api
from popsift_b200.synth import make_frame

def run(w, h, seed, setup=None, frames=2, float_img=False):
    cfg = api.Config()
    if setup:
        setup(cfg)
    img = make_frame(w, h, seed)
    if float_img:
        img = img.astype(np.float32) / np.float32(256.0)
        ps = api.PopSift(cfg, imode=api.PopSift.FloatImages, max_width=w, max_height=h, slots=1)
    else:
        ps = api.PopSift(cfg, max_width=w, max_height=h, slots=1)
    n = None
    for _ in range(frames):
        f = ps.enqueue(w, h, img).get()
        n = (f.getFeatureCount(), f.getDescriptorCount())
    ps.uninit()
    return n

only = sys.argv[1] if len(sys.argv) > 1 else "all"
cases = {
    "default": lambda: run(641, 479, 1, frames=3),
    "small": lambda: run(70, 50, 2),
    "wide": lambda: run(1500, 200, 3),
    "nonint": lambda: run(300, 220, 4, lambda c: c.setDownsampling(-0.5)),
    "ds0": lambda: run(400, 300, 5, lambda c: c.setDownsampling(0)),
    "float": lambda: run(320, 240, 6, float_img=True),
    "filter": lambda: run(640, 480, 7, lambda c: (c.setFilterMaxExtrema(100), c.setFilterGridSize(3), c.setFilterSorting("up"))),
    "direct": lambda: run(320, 240, 8, lambda c: c.setScalingMode(0)),
    "vlfeat-direct": lambda: run(320, 240, 9, lambda c: c.setGaussMode("vlfeat-direct")),
    "relative": lambda: run(320, 240, 10, lambda c: c.setGaussMode("relative")),
    "fixed9": lambda: run(320, 240, 11, lambda c: c.setGaussMode("fixed9")),
    "fixed15": lambda: run(320, 240, 12, lambda c: c.setGaussMode("fixed15")),
    "igrid": lambda: run(320, 240, 13, lambda c: c.setDescMode("igrid")),
    "notile": lambda: run(320, 240, 14, lambda c: c.setDescMode("notile")),
    "opencv": lambda: run(320, 240, 15, lambda c: c.setMode("opencv")),
}
for name, fn in cases.items():
    if only in ("all", name):
        print(name, fn(), flush=True)
if only in ("all", "match"):
    rng = np.random.default_rng(0)
    def unit(n):
        d = rng.random((n, 128), dtype=np.float32)
        return d / np.linalg.norm(d, axis=1, keepdims=True)
    l, r = unit(1100), unit(2300)
    a = api.match_descriptors(l, r, api.FeaturesDev.MATCH_TENSOR)
    b = api.match_descriptors(l, r, api.FeaturesDev.MATCH_EXACT)
    print("match rows differing", int((a != b).any(1).sum()), flush=True)
