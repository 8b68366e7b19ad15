set -u
mkdir -p gpurun_out/ab8
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short 2>&1 | grep -E "AssertionError|passed|failed|assert" | cut -c1-600
python tools/pyr_time.py 2>&1 | sed -n 2,3p
ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -c 60 --csv --log-file gpurun_out/ab8/launches.csv python tools/one_frame.py 3840 2160 5 1 > gpurun_out/ab8/ncu.log 2>&1; tail -1 gpurun_out/ab8/ncu.log
python bench.py --steps 5 --warmup 3 > gpurun_out/ab8/bench.json 2> gpurun_out/ab8/bench.err; cat gpurun_out/ab8/bench.json | head -c 300
