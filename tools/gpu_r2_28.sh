#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/summary.txt
run() { echo "=== $*" | tee -a gpurun_out/summary.txt; env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'])" | tee -a gpurun_out/summary.txt; }
run ROBOVLN_GRID_MIN=0
run ROBOVLN_GRID_MIN=24
run ROBOVLN_GRID_MIN=48
run ROBOVLN_GRID_MIN=72
run ROBOVLN_GRID_MIN=100
run ROBOVLN_GRID_MIN=0
