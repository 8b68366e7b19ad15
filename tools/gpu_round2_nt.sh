#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02nt; mkdir -p $O
cat /proc/loadavg; grep -m1 "model name" /proc/cpuinfo
timeout 200 python bench.py --steps 3 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -1 $O/bench.err
python - "$O/bench.json" <<'PY'
import json,sys
j=json.load(open(sys.argv[1]))
print("value",round(j['value']),"e2e",round(j['e2e']['value']),"pinned",round(j['e2e']['pinned_ctypes']['value']))
PY
timeout 100 python -m pytest tests -q -m gpu -x -k "cpp_dropin or popsift_demo" 2>&1 | tail -2
