#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/summary.txt
for dg in 0 96 64 48 32 16; do
  echo "=== depth grid cap=$dg" | tee -a gpurun_out/summary.txt
  ROBOVLN_DEPTH_GRID=$dg timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])" | tee -a gpurun_out/summary.txt
done
ROBOVLN_SKIP=rgb,bert ROBOVLN_DEPTH_GRID=32 timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('depth only cap32', d['ms_per_step'])" | tee -a gpurun_out/summary.txt
