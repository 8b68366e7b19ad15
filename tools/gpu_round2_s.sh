#!/usr/bin/env bash
# GPU call: debug of --gauss-mode vlfeat-direct (where do the planes differ?) + probe of the unnormalized linear float texture.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02s; mkdir -p $O
$PWD/oracle/_ref/texprobe lcoords $O/tex_lcoords.bin
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import numpy as np
import oracle_lib as ol
from popsift_b200 import api
from popsift_b200.synth import make_frame
for (w, h) in ((256, 192), (640, 480)):
    img = make_frame(w, h, 33)
    cfg = api.Config(); cfg.setMode("vlfeat"); cfg.setNormMode("classic"); cfg.setGaussMode("vlfeat-direct")
    ps = api.PopSift(cfg, max_width=w, max_height=h, slots=1)
    f = ps.enqueue(w, h, img).get()
    o = ol.Oracle(ol.make_config(mode="vlfeat", norm="classic", gauss_direct=1), w, h)
    o.run(img, 1)
    for l in range(6):
        a, b = ps.plane(0, 0, l), o.gauss(0, l)
        d = a != b
        ys, xs = np.nonzero(d)
        print(w, h, "level", l, "differing", int(d.sum()), "of", d.size, "max abs diff", float(np.abs(a - b).max()),
              "x range", (int(xs.min()), int(xs.max())) if d.any() else None, "y range", (int(ys.min()), int(ys.max())) if d.any() else None,
              "first", [(int(y), int(x), float(a[y, x]), float(b[y, x])) for y, x in list(zip(ys, xs))[:3]])
    ps.uninit(); o.close()
PY
ls -la $O
