#!/usr/bin/env bash
# GPU call: --gauss-mode fixed9 / fixed15 against the oracle and the live reference (+ reference fixtures), whole suite.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02u; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x -k "fixed or refused" > $O/pytest_fixed.txt 2>&1; tail -30 $O/pytest_fixed.txt
R=$PWD/oracle/_ref
python - <<'PY'
import sys, os
sys.path.insert(0, os.getcwd())
from popsift_b200.synth import make_frame, write_pgm
write_pgm("/tmp/f256.pgm", make_frame(256, 192, 3))
PY
for M in fixed9 fixed15; do
  rm -rf /tmp/d256 && mkdir -p /tmp/d256 && cd /tmp/d256 && $R/ref_dump -i /tmp/f256.pgm -o $OLDPWD/$O/ref_${M}_f256.bin --mode vlfeat --norm classic --gauss-mode $M --log > /dev/null 2>&1; cd $OLDPWD
  python - "$O" "$M" <<'PY'
import sys, os, hashlib, json
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
import oracle_lib as ol
O, M = sys.argv[1], sys.argv[2]
meta = {}
for oc in range(6):
    for l in range(6):
        fn = "/tmp/d256/dir-octave-dump/pyramid-o-%d-l-%d.dump" % (oc, l)
        if os.path.exists(fn):
            p = ol.read_ref_dump(fn)
            meta["g_%d_%d" % (oc, l)] = {"shape": list(p.shape), "sha256": hashlib.sha256(p.tobytes()).hexdigest()}
json.dump(meta, open(O + "/ref_%s_f256_planes.json" % M, "w"))
print(M, len(meta), "plane hashes")
PY
done
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
ls -la $O
