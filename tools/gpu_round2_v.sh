#!/usr/bin/env bash
# GPU call: --direct-scaling combined with every Gauss mode, then the whole suite.
set -uo pipefail
cd "$(dirname "$0")/.."
O=gpurun_out/r02v; mkdir -p $O
timeout 900 python -m pytest tests -q -m gpu -x -k "every_gauss_mode or refused" > $O/pytest_combo.txt 2>&1; tail -30 $O/pytest_combo.txt
timeout 1500 python -m pytest tests -q -m gpu > $O/pytest_gpu.txt 2>&1; tail -4 $O/pytest_gpu.txt
