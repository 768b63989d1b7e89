#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/summary.txt
run() { echo "=== $*" | tee -a gpurun_out/summary.txt; env "$@" timeout 600 python bench.py --steps 30 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(round(d['ms_per_step'],4), round(d['e2e']['ms_per_step'],4))" | tee -a gpurun_out/summary.txt; }
run ROBOVLN_GRID_ROUNDS=1,1,1
run ROBOVLN_GRID_ROUNDS=0,0,0
run ROBOVLN_GRID_ROUNDS=2,1,1
run ROBOVLN_GRID_ROUNDS=1,1,2
run ROBOVLN_GRID_ROUNDS=1,2,1
run ROBOVLN_GRID_ROUNDS=2,2,2
run ROBOVLN_PDL=1
run ROBOVLN_LN_FUSED=2
run ROBOVLN_ATTN=tc
