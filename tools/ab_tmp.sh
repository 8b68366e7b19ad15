set -u
mkdir -p gpurun_out/ab4
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
echo "== L0SLOTS=592"; POPSIFT_B200_L0SLOTS=592 python tools/pyr_time.py 2>&1 | sed -n 2,3p
echo "== L0SLOTS=740"; POPSIFT_B200_L0SLOTS=740 python tools/pyr_time.py 2>&1 | sed -n 2,3p
echo "== L0SLOTS=888"; python tools/pyr_time.py 2>&1 | sed -n 2,9p
echo "== L0SLOTS=1184"; POPSIFT_B200_L0SLOTS=1184 python tools/pyr_time.py 2>&1 | sed -n 2,3p
ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_active.min,sm__cycles_elapsed.avg,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:march_level0 -c 1 --csv --log-file gpurun_out/ab4/l0.csv python tools/one_frame.py 3840 2160 5 1 > gpurun_out/ab4/ncu.log 2>&1; tail -2 gpurun_out/ab4/ncu.log
