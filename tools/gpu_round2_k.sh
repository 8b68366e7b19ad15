#!/usr/bin/env bash
# GPU call: matcher epilogue (branch-free group minimum, TMEM / |b|^2 prefetch), enqueue copy A/B.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02k; mkdir -p $O
( cat /proc/loadavg; nproc ) > $O/host.txt 2>&1; cat $O/host.txt
timeout 900 python -m pytest tests -q -m gpu -x -k "match" > $O/pytest_match.txt 2>&1; tail -4 $O/pytest_match.txt
for C in 128 64 128r 64r; do
  POPSIFT_B200_MATCH_RING=$C timeout 300 python tools/match_bench.py $O/match_bench_$C.json 2> $O/match_bench_$C.err | python -c "
import json,sys
j=json.loads(sys.stdin.read()); print('$C', 'tensor_ms', round(j['tensor_ms'],3), 'exact_ms', round(j['exact_ms'],2), 'differing', j['rows_differing_tensor_vs_exact'], 'frac', round(j['roofline']['frac'],3))"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:match_tc_kernel -c 1 -o /tmp/mt python tools/match_bench.py > $O/ncu_match.log 2>&1; tail -2 $O/ncu_match.log
ncu -i /tmp/mt.ncu-rep --page raw --csv > $O/match_raw.csv 2>/dev/null
ncu -i /tmp/mt.ncu-rep --page source --csv > $O/match_source.csv 2>/dev/null
python tools/ncu_summary.py $O/match_raw.csv > $O/match_summary.tsv 2>&1; cat $O/match_summary.tsv
timeout 900 python tools/e2e_ab.py $O/e2e_ab.json 4 > $O/e2e_ab.txt 2>&1; cat $O/e2e_ab.txt
timeout 1500 python -m pytest tests -q -m gpu -x > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
du -sh gpurun_out
