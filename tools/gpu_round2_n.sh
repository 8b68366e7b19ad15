#!/usr/bin/env bash
# GPU call: level kernels without CTA barriers in interior strips (warp-owned ring columns) vs with (POPSIFT_B200_OWNED=0).
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02n; mkdir -p $O
( cat /proc/loadavg; nproc ) > $O/host.txt 2>&1; cat $O/host.txt
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
for W in 1 0 1 0; do
  POPSIFT_B200_OWNED=$W timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_own$W.json 2> $O/bench_own$W.err; tail -2 $O/bench_own$W.err
  python - "$O/bench_own$W.json" <<'PY'
import json,sys
j=json.load(open(sys.argv[1]))
print(sys.argv[1],"value",round(j['value']),"e2e",round(j['e2e']['value']),"pinned",round(j['e2e']['pinned_ctypes']['value']),"roofline",round(j['roofline']['frac'],4),"ms",round(j['roofline']['ms'],4),"dominant",round(j['roofline']['dominant_kernel']['frac'],4))
PY
done
for W in 1 0; do
  POPSIFT_B200_OWNED=$W ncu --metrics gpu__time_duration.sum --clock-control none -c 45 --csv --log-file $O/launches_own$W.csv python tools/one_frame.py 3840 2160 5 1 > $O/one_frame_own$W.log 2>&1
  python tools/summarize_launches.py $O/launches_own$W.csv > $O/launches_own$W.txt 2>&1; head -8 $O/launches_own$W.txt
done
du -sh gpurun_out
