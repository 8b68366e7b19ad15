set -u
O=gpurun_out/ab17; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short 2>&1 | grep -E "AssertionError|passed|failed|assert|Error" | cut -c1-600
python tools/pyr_time.py 2>&1 | sed -n 2,3p
ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_elapsed.avg --clock-control none -c 60 --csv --log-file $O/launches.csv python tools/one_frame.py 3840 2160 5 1 > $O/ncu.log 2>&1; tail -1 $O/ncu.log
python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; cat $O/bench.json | head -c 300
