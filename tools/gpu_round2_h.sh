#!/usr/bin/env bash
# GPU call: level-0 coordinate tests against the live reference, whole GPU suite, ncu --set full of one 4K frame (raw page
# as CSV + source page of the octave-0 level kernels), bench.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02h; mkdir -p $O
( cat /proc/loadavg; nproc ) > $O/host.txt 2>&1; cat $O/host.txt
timeout 900 python -m pytest tests -q -m gpu -x -k "level0_planes or live_against" > $O/pytest_level0.txt 2>&1; tail -15 $O/pytest_level0.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -6 $O/pytest_gpu.txt
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; tail -c 600 $O/bench.json; tail -2 $O/bench.err
# full-set capture of one frame
timeout 900 ncu --set full --clock-control none --import-source on -c 45 -o /tmp/full_frame python tools/one_frame.py 3840 2160 5 1 > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
ncu -i /tmp/full_frame.ncu-rep --page raw --csv > $O/full_frame_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/full_frame_raw.csv > $O/full_frame_summary.tsv 2> $O/ncu_summary.err; head -50 $O/full_frame_summary.tsv
# source pages (per-instruction stall samples) of two level kernels of octave 0: the 4th and 5th march_level launch (R=10, R=13)
for K in 3 4; do
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:march_level_kernel -s $K -c 1 -o /tmp/lv$K python tools/one_frame.py 3840 2160 5 1 > $O/ncu_lv$K.log 2>&1
  ncu -i /tmp/lv$K.ncu-rep --page source --csv > $O/source_lv$K.csv 2>/dev/null
  if [ $(stat -c %s /tmp/lv$K.ncu-rep) -lt 12000000 ]; then cp /tmp/lv$K.ncu-rep $O/; fi
done
ls -la $O | head -30
du -sh gpurun_out
