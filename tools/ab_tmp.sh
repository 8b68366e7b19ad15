set -u
O=gpurun_out/ab19; mkdir -p $O
python -m pytest tests/test_gpu_parity.py -m gpu -q --tb=short 2>&1 | grep -E "AssertionError|passed|failed|assert|Error" | cut -c1-600
for S in 4 6 8; do
POPSIFT_BENCH_SLOTS=$S python bench.py --steps 5 --warmup 3 > $O/bench_s$S.json 2> $O/bench.err; python - <<PY
import json
j=json.load(open("$O/bench_s$S.json")); print("slots $S value", round(j["value"]), "e2e", round(j["e2e"]["value"]), "frac", round(j["roofline"]["frac"],3))
PY
done
ncu --metrics gpu__time_duration.sum,sm__inst_executed.sum,smsp__issue_active.avg.pct_of_peak_sustained_active -k regex:descriptor --clock-control none -c 1 --csv --log-file $O/desc.csv python tools/one_frame.py 3840 2160 5 1 > $O/ncu.log 2>&1; grep -E "duration|inst_executed" $O/desc.csv | cut -d, -f13-
