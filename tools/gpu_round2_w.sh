#!/usr/bin/env bash
# GPU call: images in flight per GPU (slots) 4 / 6 / 8.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02w; mkdir -p $O
cat /proc/loadavg
for S in 4 6 8 4 6; do
  POPSIFT_BENCH_SLOTS=$S timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_s$S.json 2> $O/bench_s$S.err; tail -1 $O/bench_s$S.err
  python - "$O/bench_s$S.json" $S <<'PY'
import json,sys
j=json.load(open(sys.argv[1]))
print("slots",sys.argv[2],"value",round(j['value']),"e2e",round(j['e2e']['value']),"pinned",round(j['e2e']['pinned_ctypes']['value']))
PY
done
cat /proc/loadavg
