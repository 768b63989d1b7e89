#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/summary.txt
for er in 1,1,1 2,1,1 1,1,2 2,1,2 1,0,1 1,2,1 2,0,2 3,1,2; do
  echo "=== rounds rgb,depth,bert=$er" | tee -a gpurun_out/summary.txt
  ROBOVLN_GRID_ROUNDS=$er timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])" | tee -a gpurun_out/summary.txt
done
