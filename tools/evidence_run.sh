#!/usr/bin/env bash
# Round evidence on one B200: full-set ncu capture of the kernels of one 4K frame (exported as CSV on the box;
# the .ncu-rep is kept only when it is small), launch list, bench (both arms).
set -u
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-evidence}; mkdir -p $O
ncu --set full --clock-control none -c 40 -o /tmp/full_frame python tools/one_frame.py 3840 2160 5 1 > $O/ncu_full.log 2>&1; tail -1 $O/ncu_full.log
ncu -i /tmp/full_frame.ncu-rep --page raw --csv > $O/full_frame_raw.csv 2>/dev/null
ls -la /tmp/full_frame.ncu-rep | awk '{print "ncu-rep bytes:", $5}'
if [ $(stat -c %s /tmp/full_frame.ncu-rep) -lt 30000000 ]; then cp /tmp/full_frame.ncu-rep $O/; fi
if [ "${LIST:-1}" = 1 ]; then
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv python tools/one_frame.py 3840 2160 5 > $O/ncu_launches.log 2>&1
fi
if [ "${BENCH:-1}" = 1 ]; then
python bench.py --steps ${STEPS:-5} --warmup 3 > $O/bench.json 2> $O/bench.err; tail -c 300 $O/bench.json; tail -2 $O/bench.err
python bench.py --impl reference --steps ${STEPS:-5} --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 300 $O/bench_ref.json
fi
du -sh gpurun_out
