#!/usr/bin/env bash
# GPU call: CUDA-graph replay of the per-frame pipeline; host-load diagnostics; texture coordinate probe at non-integer scale;
# matcher timing vs the reference's compute_distance.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02g; mkdir -p $O
R=$PWD/oracle/_ref
( cat /proc/loadavg; nproc; uptime ) > $O/host.txt 2>&1; cat $O/host.txt
$R/texprobe coords 700 500 0.5 $O/tex_c700_up05.bin
$R/texprobe coords 640 480 1.5 $O/tex_c640_up15.bin
timeout 1500 python -m pytest tests -q -m gpu -s > $O/pytest_gpu.txt 2>&1; tail -6 $O/pytest_gpu.txt
for g in 1 0; do
  POPSIFT_B200_GRAPH=$g timeout 600 python bench.py --steps 5 --warmup 3 > $O/bench_graph$g.json 2> $O/bench_graph$g.err
  python - "$O/bench_graph$g.json" <<'PY'
import json,sys
try:
    j=json.load(open(sys.argv[1]))
    print(sys.argv[1],"value",round(j['value']),"e2e",round(j['e2e']['value']),"pinned",round(j['e2e']['pinned_ctypes']['value']),"roofline",round(j['roofline']['frac'],3),"ms/step",round(j['ms_per_step'],2))
except Exception as e: print("bench failed", e)
PY
done
cat /proc/loadavg
timeout 600 python tools/match_bench.py $O/match_bench.json --pgm $O/mpgm 2> $O/match_bench.err | tail -1
ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/ref_match_launches.csv -k regex:compute_distance $R/ref_dump -i $O/mpgm/m0.pgm -i $O/mpgm/m1.pgm --octaves 5 --match > /dev/null 2> $O/ref_match.err
grep compute_distance $O/ref_match_launches.csv | tail -2
rm -rf $O/mpgm
ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $O/launches.csv python tools/one_frame.py 3840 2160 5 2 > $O/one_frame.log 2>&1
python tools/summarize_launches.py $O/launches.csv > $O/launches.txt 2>&1; head -8 $O/launches.txt
