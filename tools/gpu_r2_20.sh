#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/summary.txt
for sk in none depth rgb bert depth,rgb depth,bert rgb,bert; do
  echo "=== skip=$sk" | tee -a gpurun_out/summary.txt
  ROBOVLN_SKIP=$sk timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])" | tee -a gpurun_out/summary.txt
done
