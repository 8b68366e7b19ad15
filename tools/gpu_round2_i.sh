#!/usr/bin/env bash
# GPU call: block-wise ring addressing in the level kernels (RING multiple of 8); orientation batch A/B; orientation experiment.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02i; mkdir -p $O
( cat /proc/loadavg; nproc ) > $O/host.txt 2>&1; cat $O/host.txt
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_sel.txt 2>&1; tail -5 $O/pytest_sel.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -2 $O/bench.err
python - "$O/bench.json" <<'PY'
import json,sys
j=json.load(open(sys.argv[1]))
print("value",round(j['value']),"e2e",round(j['e2e']['value']),"roofline",round(j['roofline']['frac'],4), j['roofline'].get('ms'), "dominant", j['roofline'].get('dominant_kernel',{}).get('frac'))
PY
for B in 1 2 4; do
  POPSIFT_B200_ORI_BATCH=$B ncu --metrics gpu__time_duration.sum --clock-control none -c 45 --csv --log-file $O/launches_b$B.csv python tools/one_frame.py 3840 2160 5 1 > $O/one_frame_b$B.log 2>&1
  echo "batch $B: $(grep orientation_kernel $O/launches_b$B.csv | tail -1 | awk -F'","' '{print $NF}')"
done
python tools/summarize_launches.py $O/launches_b1.csv > $O/launches.txt 2>&1; head -40 $O/launches.txt
for B in 2 4; do
  POPSIFT_B200_ORI_BATCH=$B timeout 900 python -m pytest tests -q -m gpu -x -k "benchmark_workload or features_vs_golden" > $O/pytest_batch$B.txt 2>&1; tail -2 $O/pytest_batch$B.txt
done
timeout 600 python tools/match_bench.py $O/match_bench.json 2> $O/match_bench.err | tail -1
timeout 900 python tools/ori_experiment.py $O/ori_experiment.json > $O/ori_experiment.txt 2>&1; cat $O/ori_experiment.txt | tail -12
du -sh gpurun_out
