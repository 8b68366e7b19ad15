#!/usr/bin/env bash
# GPU call: pair-wise stage 1 of the extrema kernel, guard-ring descriptor scatter, descriptor modes, opencv gauss mode.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02f; mkdir -p $O
timeout 1500 python -m pytest tests -q -m gpu -s > $O/pytest_gpu.txt 2>&1; tail -12 $O/pytest_gpu.txt
cp gpurun_out/bench32_angle_diffs.json $O/ 2>/dev/null
timeout 300 python tools/pyr_time.py > $O/pyr_time.txt 2>&1; head -3 $O/pyr_time.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
try:
    j=json.load(open('gpurun_out/r02f/bench_n1.json'))
    print("value",j['value'],"e2e",j['e2e']['value'],"roofline",j['roofline']['frac'],j['roofline']['ms'],"dom",j['roofline']['dominant_kernel']['frac'])
except Exception as e: print("bench failed", e)
PY
tail -3 $O/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python tools/one_frame.py 3840 2160 5 2 > $O/one_frame.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches.txt 2>&1; head -16 $O/launches.txt
