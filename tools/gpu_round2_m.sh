#!/usr/bin/env bash
# GPU call: extrema kernel (next-region prefetch; 3 vs 4 resident CTAs), matcher with its own memory pool, full tests, bench.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02m; mkdir -p $O
( cat /proc/loadavg; nproc ) > $O/host.txt 2>&1; cat $O/host.txt
for C in 4 3; do
  POPSIFT_B200_EXTREMA_CTAS=$C ncu --metrics gpu__time_duration.sum --clock-control none -c 45 --csv --log-file $O/launches_c$C.csv python tools/one_frame.py 3840 2160 5 1 > $O/one_frame_c$C.log 2>&1
  echo "extrema CTAs/SM $C: $(grep cand_extrema $O/launches_c$C.csv | tail -1 | awk -F'","' '{print $NF}') ns; $(tail -1 $O/one_frame_c$C.log)"
done
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
POPSIFT_B200_EXTREMA_CTAS=3 timeout 900 python -m pytest tests -q -m gpu -x -k "benchmark_workload or features_vs_golden or planes_and_extrema or edge" > $O/pytest_c3.txt 2>&1; tail -2 $O/pytest_c3.txt
timeout 300 python tools/match_bench.py $O/match_bench.json 2> $O/match_bench.err | tail -1
for C in 4 3; do
  POPSIFT_B200_EXTREMA_CTAS=$C timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_c$C.json 2> $O/bench_c$C.err; tail -2 $O/bench_c$C.err
  python - "$O/bench_c$C.json" <<'PY'
import json,sys
j=json.load(open(sys.argv[1]))
print(sys.argv[1],"value",round(j['value']),"e2e",round(j['e2e']['value']),"pinned",round(j['e2e']['pinned_ctypes']['value']),"roofline",round(j['roofline']['frac'],4))
PY
done
cat /proc/loadavg
du -sh gpurun_out
