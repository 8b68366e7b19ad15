#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02z; mkdir -p $O
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_run.py default > $O/initcheck_default.txt 2>&1; echo "initcheck rc $?"
grep -A3 "Uninitialized __global__" $O/initcheck_default.txt | grep -E "at " | sed 's/(.*//' | sort | uniq -c | sort -rn | head; grep "ERROR SUMMARY" $O/initcheck_default.txt
timeout 600 compute-sanitizer --tool initcheck --error-exitcode 9 python tools/sanitize_run.py match > $O/initcheck_match.txt 2>&1; grep "ERROR SUMMARY" $O/initcheck_match.txt; grep -A3 "Uninitialized __global__" $O/initcheck_match.txt | grep -E "at " | sed 's/(.*//' | sort | uniq -c | sort -rn | head -5
