"""Runs N frames of WxH through the public API (profiling target for ncu)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from popsift_b200 import api
from popsift_b200.synth import make_frame
w, h = int(sys.argv[1]), int(sys.argv[2])
octv = int(sys.argv[3]) if len(sys.argv) > 3 else -1
n = int(sys.argv[4]) if len(sys.argv) > 4 else 2
cfg = api.Config()
if octv > 0:
    cfg.setOctaves(octv)
ps = api.PopSift(cfg, max_width=w, max_height=h, slots=1)
img = make_frame(w, h, 7)
for _ in range(n):
    f = ps.enqueue(w, h, img).get()
print(f.getFeatureCount(), f.getDescriptorCount())
ps.uninit()
