#!/usr/bin/env bash
# One GPU-box session: tests, smoke, bench (both arms), ncu launch list.  Outputs in gpurun_out/$1/
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/${1:-round}; mkdir -p $O
( python -m pytest tests -m gpu -x -q 2>&1 | tail -15 ) > $O/pytest_gpu.txt
cat $O/pytest_gpu.txt | tail -5
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.txt 2>&1; tail -2 $O/smoke.txt
python bench.py --steps ${STEPS:-5} --warmup 3 > $O/bench.json 2> $O/bench.err; tail -c 3000 $O/bench.json; tail -3 $O/bench.err
python bench.py --impl reference --steps ${STEPS:-5} --warmup 2 > $O/bench_ref.json 2> $O/bench_ref.err; tail -c 1500 $O/bench_ref.json
if [ "${NCU:-1}" = 1 ]; then
  ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches.csv \
      python tools/one_frame.py 3840 2160 5 > $O/ncu_launches.log 2>&1
  python tools/summarize_launches.py $O/launches.csv > $O/launches_summary.txt 2>&1; cat $O/launches_summary.txt
fi
