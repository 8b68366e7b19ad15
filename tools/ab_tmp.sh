set -u
mkdir -p gpurun_out/ab2
echo "== OLD"; POPSIFT_B200_LIB=$PWD/popsift_b200/lib_old/libpopsift_b200.so python tools/pyr_time.py 2>&1 | sed -n 2p
echo "== NEW"; python tools/pyr_time.py 2>&1 | tail -13
echo "== OLD again"; POPSIFT_B200_LIB=$PWD/popsift_b200/lib_old/libpopsift_b200.so python tools/pyr_time.py 2>&1 | sed -n 2p
echo "== NEW again"; python tools/pyr_time.py 2>&1 | sed -n 2p
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
ncu --metrics gpu__time_duration.sum,sm__cycles_active.avg,sm__cycles_active.max,sm__cycles_active.min,sm__cycles_elapsed.avg,sm__inst_executed.sum --clock-control none -k regex:march_level -c 12 --csv --log-file gpurun_out/ab2/bal.csv python tools/one_frame.py 3840 2160 5 1 > gpurun_out/ab2/ncu.log 2>&1; tail -2 gpurun_out/ab2/ncu.log
