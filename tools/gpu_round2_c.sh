#!/usr/bin/env bash
# GPU call: TMA-staged pyramid kernels -- memcheck on a small frame, the GPU test suite, pyramid timing with and
# without programmatic dependent launch, orientation parity in both accumulation modes, a short bench, a launch list.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02c; mkdir -p $O
timeout 300 compute-sanitizer --tool memcheck --print-limit 20 python tools/one_frame.py 640 480 -1 1 > $O/memcheck.txt 2>&1; tail -4 $O/memcheck.txt
timeout 900 python -m pytest tests -x -q -m gpu -s > $O/pytest_gpu.txt 2>&1; tail -8 $O/pytest_gpu.txt
POPSIFT_B200_ORI_LANESUM=1 timeout 600 python -m pytest tests -x -q -m gpu -s -k "benchmark_workload" > $O/pytest_lanesum.txt 2>&1; tail -5 $O/pytest_lanesum.txt
for pdl in 1 0; do
  POPSIFT_B200_PDL=$pdl timeout 300 python tools/pyr_time.py > $O/pyr_time_pdl$pdl.txt 2>&1; head -3 $O/pyr_time_pdl$pdl.txt
done
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; python - <<'PY'
import json
j=json.load(open('gpurun_out/r02c/bench_n1.json'))
print("value",j['value'],"e2e",j['e2e']['value'],"roofline",j['roofline']['frac'],j['roofline']['ms'],"dom",j['roofline']['dominant_kernel']['frac'])
PY
tail -3 $O/bench_n1.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python tools/one_frame.py 3840 2160 5 2 > $O/one_frame.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches.txt 2>&1; head -16 $O/launches.txt
