#!/usr/bin/env bash
# GPU call: grid-filter reference runs with limits that trigger the filter, the 8-bit texture at general
# fractions, the GPU test suite, a short bench of both arms, and a launch list of one 4K frame.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02b; mkdir -p $O
R=$PWD/oracle/_ref
python -m popsift_b200.synth 640 480 1 $O/f640.pgm
python -m popsift_b200.synth 1280 960 5 $O/f1280.pgm
$R/texprobe upairs $O/tex_upairs.bin
for fm in 100 200; do for g in 2 3; do for s in up down random; do
  $R/ref_dump -i $O/f640.pgm -o $O/f640_filter_${fm}_${g}_${s}.bin --mode vlfeat --norm classic --filter-max-extrema $fm --filter-grid $g --filter-sort $s 2>&1 | tail -1
done; done; done
for s in up down; do
  $R/ref_dump -i $O/f1280.pgm -o $O/f1280_filter_1000_4_${s}.bin --filter-max-extrema 1000 --filter-grid 4 --filter-sort $s 2>&1 | tail -1
done
$R/ref_dump -i $O/f1280.pgm -o $O/f1280_nofilter.bin 2>&1 | tail -1
rm -f $O/*.pgm
python -m pytest tests -x -q -m gpu -s > $O/pytest_gpu.txt 2>&1; tail -15 $O/pytest_gpu.txt
python bench.py --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -c 1500 $O/bench_n1.json; tail -3 $O/bench_n1.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/bench_ref_n1.json 2> $O/bench_ref_n1.err; tail -c 600 $O/bench_ref_n1.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python tools/one_frame.py 3840 2160 5 2 > $O/one_frame.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches.txt 2>&1; tail -30 $O/launches.txt
