#!/usr/bin/env python
"""A/B of the enqueue copy (POPSIFT_B200_COPY_THREADS=1 vs 4) through popsift_b200/bin/api_bench, alternating runs so that
host-load drift hits both arms alike.   python tools/e2e_ab.py OUT.json [rounds]"""
import json, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench
def main(out, rounds):
    frames = bench.synth_frames(bench.FRAMES_PER_STEP, 0)
    exe = os.path.join(ROOT, "popsift_b200", "bin", "api_bench")
    res = {"1": [], "4": [], "loadavg": []}
    with tempfile.TemporaryDirectory() as td:
        cmd = [exe, "--octaves", str(bench.OCTAVES), "--levels", str(bench.LEVELS), "--slots", str(bench.SLOTS), "--bench", "3", "2"]
        for p in bench.write_frames(frames, td):
            cmd += ["-i", p]
        for r in range(rounds):
            for t in ("1", "4"):
                o = subprocess.run(cmd, env={**os.environ, "POPSIFT_B200_COPY_THREADS": t}, capture_output=True, text=True, check=True)
                j = json.loads([l for l in o.stdout.splitlines() if l.startswith("{")][-1])
                res[t].append(j["mpix_per_s"])
            res["loadavg"].append(open("/proc/loadavg").read().split()[0])
            print(r, res["1"][-1], res["4"][-1], res["loadavg"][-1], flush=True)
    json.dump(res, open(out, "w"), indent=1)
if __name__ == "__main__":
    main(sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 4)
