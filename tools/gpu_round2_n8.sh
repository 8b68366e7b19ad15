#!/usr/bin/env bash
# 8-GPU box: our arm at N = 8 (frames sharded, no collective; ranks on both NUMA nodes)
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02n8; mkdir -p $O
( cat /proc/loadavg; nvidia-smi -L | wc -l; nvidia-smi topo -m | grep -E "^GPU[0-7]" | awk '{print $1, $(NF-2), $(NF-1)}' ) > $O/host.txt 2>&1; cat $O/host.txt
timeout 200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 bench.py --gpus 8 --steps 3 --warmup 3 > $O/bench_n8.json 2> $O/bench_n8.err; echo "rc $?"; tail -2 $O/bench_n8.err | cut -c1-300
python - "$O/bench_n8.json" <<'PY'
import json, sys
for l in reversed(open(sys.argv[1]).read().splitlines()):
    if l.startswith("{"):
        j = json.loads(l)
        print("N=8 value %.0f e2e %.0f pinned %.0f n_gpus %d numa %s" % (j["value"], j["e2e"]["value"], j["e2e"]["pinned_ctypes"]["value"], j["n_gpus"], j.get("numa")))
        break
else:
    print("no JSON line")
PY
