#!/usr/bin/env bash
# Round-2 evidence on one B200: GPU test suite, smoke, bench (both arms), launch lists, ncu --set full of one 4K frame and of
# the matcher, matcher timing.  Outputs in gpurun_out/r02final/ (copied into profiles/ by hand).
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02final; mkdir -p $O
( cat /proc/loadavg; nproc; nvidia-smi --query-gpu=name,driver_version,clocks.max.sm --format=csv,noheader ) > $O/host.txt 2>&1; cat $O/host.txt
timeout 1500 python -m pytest tests -q -m gpu -s > $O/pytest_gpu_full.txt 2>&1; tail -3 $O/pytest_gpu_full.txt > $O/pytest_gpu.txt; grep -h "bench32 parity\|desc mode" $O/pytest_gpu_full.txt >> $O/pytest_gpu.txt; cat $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
timeout 900 python bench.py --steps 10 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 400 $O/bench_n1.json; tail -2 $O/bench_n1.err
timeout 900 python bench.py --impl reference --steps 10 --warmup 2 > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err; tail -c 600 $O/bench_reference_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_4k.csv python tools/one_frame.py 3840 2160 5 > $O/one_frame.log 2>&1
python tools/summarize_launches.py $O/launches_4k.csv > $O/ncu_launches_4k.txt 2>&1; cat $O/ncu_launches_4k.txt
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_bench.csv python bench.py --steps 1 --warmup 3 > $O/bench_under_ncu.log 2>&1
python tools/summarize_launches.py $O/launches_bench.csv > $O/ncu_launches_bench.txt 2>&1; head -14 $O/ncu_launches_bench.txt
timeout 900 ncu --set full --clock-control none --import-source on -c 45 -o /tmp/full_frame python tools/one_frame.py 3840 2160 5 1 > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
ncu -i /tmp/full_frame.ncu-rep --page raw --csv > $O/ncu_full_frame_4k.csv 2>/dev/null
python tools/ncu_summary.py $O/ncu_full_frame_4k.csv > $O/ncu_full_summary.tsv 2> $O/ncu_summary.err
timeout 300 python tools/match_bench.py $O/match_bench.json 2> $O/match_bench.err | tail -1
timeout 600 ncu --set full --clock-control none -k regex:match_ -c 5 -o /tmp/mt python tools/match_bench.py > $O/ncu_match.log 2>&1
ncu -i /tmp/mt.ncu-rep --page raw --csv > $O/ncu_match_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/ncu_match_raw.csv > $O/ncu_match_summary.tsv 2>&1; cat $O/ncu_match_summary.tsv
ls -la $O; du -sh gpurun_out
