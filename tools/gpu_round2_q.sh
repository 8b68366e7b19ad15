#!/usr/bin/env bash
# 2-GPU box: N = 1 and N = 2 of both bench arms back to back (weak scaling: frames sharded across ranks, no collective).
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02q; mkdir -p $O
( cat /proc/loadavg; nproc; nvidia-smi -L; nvidia-smi topo -m | head -6 ) > $O/host.txt 2>&1; cat $O/host.txt
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 > $O/bench_n1.json 2> $O/bench_n1.err; tail -2 $O/bench_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > $O/bench_n2.json 2> $O/bench_n2.err; tail -3 $O/bench_n2.err
timeout 600 python bench.py --impl reference --gpus 1 --steps 5 --warmup 2 > $O/bench_reference_n1.json 2> $O/bench_reference_n1.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 > $O/bench_reference_n2.json 2> $O/bench_reference_n2.err; tail -3 $O/bench_reference_n2.err
python - "$O" <<'PY'
import json, sys
O = sys.argv[1]
def last_json(fn):
    for l in reversed(open(fn).read().splitlines()):
        if l.startswith("{"):
            return json.loads(l)
a1, a2 = last_json(O + "/bench_n1.json"), last_json(O + "/bench_n2.json")
r1, r2 = last_json(O + "/bench_reference_n1.json"), last_json(O + "/bench_reference_n2.json")
print("ours  N=1 value %.0f e2e %.0f | N=2 value %.0f e2e %.0f | scaling value %.3f e2e %.3f" % (a1["value"], a1["e2e"]["value"], a2["value"], a2["e2e"]["value"], a2["value"] / (2 * a1["value"]), a2["e2e"]["value"] / (2 * a1["e2e"]["value"])))
print("ref   N=1 %.0f | N=2 %.0f | scaling %.3f" % (r1["value"], r2["value"], r2["value"] / (2 * r1["value"])))
print("numa", a2.get("numa"), a2.get("clocks"))
PY
