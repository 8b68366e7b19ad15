#!/usr/bin/env bash
# GPU call: orientation sq_dist contraction fix (angle parity), parallel host copy (e2e), matcher ring A/B + ncu.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02j; mkdir -p $O
( cat /proc/loadavg; nproc ) > $O/host.txt 2>&1; cat $O/host.txt
timeout 1500 python -m pytest tests -q -m gpu -x -s > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt; grep -h "bench32 parity" $O/pytest_gpu.txt
timeout 900 python tools/ori_experiment.py $O/ori_experiment.json > $O/ori_experiment.txt 2>&1; tail -9 $O/ori_experiment.txt
for T in 4 1; do
  POPSIFT_B200_COPY_THREADS=$T timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_copy$T.json 2> $O/bench_copy$T.err; tail -2 $O/bench_copy$T.err
  python - "$O/bench_copy$T.json" <<'PY'
import json,sys
j=json.load(open(sys.argv[1]))
print(sys.argv[1],"value",round(j['value']),"e2e",round(j['e2e']['value']),"pinned",round(j['e2e']['pinned_ctypes']['value']),"roofline",round(j['roofline']['frac'],4))
PY
done
cat /proc/loadavg
for C in 128 64 128r 64r; do
  POPSIFT_B200_MATCH_RING=$C timeout 300 python tools/match_bench.py $O/match_bench_$C.json 2> $O/match_bench_$C.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$C', 'tensor_ms', round(j['tensor_ms'],3), 'exact_ms', round(j['exact_ms'],2), 'differing', j['rows_differing_tensor_vs_exact'], 'frac', round(j['roofline']['frac'],3))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_tc_kernel -c 1 -o /tmp/mt python tools/match_bench.py > $O/ncu_match.log 2>&1; tail -2 $O/ncu_match.log
ncu -i /tmp/mt.ncu-rep --page raw --csv > $O/match_raw.csv 2>/dev/null
ncu -i /tmp/mt.ncu-rep --page source --csv > $O/match_source.csv 2>/dev/null
python tools/ncu_summary.py $O/match_raw.csv > $O/match_summary.tsv 2>&1; cat $O/match_summary.tsv
ncu --metrics gpu__time_duration.sum --clock-control none -c 45 --csv --log-file $O/launches.csv python tools/one_frame.py 3840 2160 5 1 > $O/one_frame.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches.txt 2>&1; head -16 $O/launches.txt
du -sh gpurun_out
