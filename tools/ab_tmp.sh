set -u
O=gpurun_out/ab21; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short 2>&1 | grep -E "AssertionError|passed|failed|assert|Error" | cut -c1-600
ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared_op_atom.sum,smsp__inst_executed_op_shared_atom.sum -k regex:descriptor --clock-control none -c 1 --csv --log-file $O/desc.csv python tools/one_frame.py 3840 2160 5 1 > $O/ncu.log 2>&1; grep -E "duration|inst_executed|wavefronts|issue" $O/desc.csv | cut -d, -f13-
python bench.py --steps 5 --warmup 3 > $O/bench.json 2> $O/bench.err; python - <<PY
import json
j=json.load(open("$O/bench.json")); print("value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "frac", round(j["roofline"]["frac"],3))
PY
