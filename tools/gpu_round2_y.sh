#!/usr/bin/env bash
# compute-sanitizer over small workloads of every kernel family
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02y; mkdir -p $O
timeout 1500 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitize_run.py > $O/memcheck.txt 2>&1; echo "memcheck rc $?"; grep -E "ERROR SUMMARY|Invalid|out of bounds|rows differing" $O/memcheck.txt | head -20; tail -3 $O/memcheck.txt
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py default > $O/racecheck_default.txt 2>&1; echo "racecheck rc $?"; grep -E "RACECHECK SUMMARY|hazard" $O/racecheck_default.txt | head -10
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitize_run.py match > $O/racecheck_match.txt 2>&1; echo "racecheck match rc $?"; grep -E "RACECHECK SUMMARY|hazard" $O/racecheck_match.txt | head -10
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_run.py default > $O/initcheck_default.txt 2>&1; echo "initcheck rc $?"; grep -E "ERROR SUMMARY|Uninitialized" $O/initcheck_default.txt | head -10
