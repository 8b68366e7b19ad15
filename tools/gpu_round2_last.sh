#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02last; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -3 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -1 $O/smoke.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -1 $O/bench.err
python - "$O/bench.json" <<'PY'
import json,sys
j=json.load(open(sys.argv[1]))
print("value",round(j['value']),"e2e",round(j['e2e']['value']),"pinned",round(j['e2e']['pinned_ctypes']['value']),"roofline",round(j['roofline']['frac'],4))
PY
