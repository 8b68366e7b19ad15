set -u
mkdir -p gpurun_out/ab3
echo "== UNIFORM"; POPSIFT_B200_UNIFORM=1 python tools/pyr_time.py 2>&1 | tail -12
echo "== LIGHT/HEAVY"; python tools/pyr_time.py 2>&1 | tail -12
echo "== UNIFORM again"; POPSIFT_B200_UNIFORM=1 python tools/pyr_time.py 2>&1 | sed -n 2p
echo "== LIGHT/HEAVY again"; python tools/pyr_time.py 2>&1 | sed -n 2p
echo "== 1080p uniform";  POPSIFT_B200_UNIFORM=1 python tools/pyr_time.py 1920 1080 5 2>&1 | sed -n 2p
echo "== 1080p l/h";  python tools/pyr_time.py 1920 1080 5 2>&1 | sed -n 2p
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3
