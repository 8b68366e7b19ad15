#!/usr/bin/env bash
# GPU call: source-level profile of the extrema kernel; kernel list of one ps_match call.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02l; mkdir -p $O
( cat /proc/loadavg; nproc ) > $O/host.txt 2>&1; cat $O/host.txt
timeout 300 ncu --set full --clock-control none --import-source on -k regex:cand_extrema -c 1 -o /tmp/ex python tools/one_frame.py 3840 2160 5 1 > $O/ncu_ex.log 2>&1; tail -1 $O/ncu_ex.log
ncu -i /tmp/ex.ncu-rep --page source --csv > $O/extrema_source.csv 2>/dev/null
ncu -i /tmp/ex.ncu-rep --page raw --csv > $O/extrema_raw.csv 2>/dev/null
python tools/ncu_summary.py $O/extrema_raw.csv
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:match_ --csv --log-file $O/match_launches.csv python tools/match_bench.py > $O/match_list.log 2>&1
python tools/summarize_launches.py $O/match_launches.csv | head; grep -c match_tc $O/match_launches.csv
tail -12 $O/match_launches.csv | awk -F'","' '{print $5, $NF}'
python - <<'PY'
# candidate statistics of one 4K frame: pairs per region
import sys, os
sys.path.insert(0, os.getcwd())
import numpy as np
from popsift_b200 import api
from popsift_b200.synth import make_frame
cfg = api.Config(); cfg.setOctaves(5)
ps = api.PopSift(cfg, max_width=3840, max_height=2160, slots=1)
f = ps.enqueue(3840, 2160, make_frame(3840, 2160, 7)).get()
print("features", f.getFeatureCount())
ps.uninit()
PY
