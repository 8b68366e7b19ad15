#!/usr/bin/env bash
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02x; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x -k "other_pyramid_modes" > $O/pytest_new.txt 2>&1; tail -12 $O/pytest_new.txt
