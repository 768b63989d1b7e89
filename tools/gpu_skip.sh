#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
for sk in none depth bert rgb depth,bert rgb,bert rgb,depth; do
  echo "== skip $sk"
  ROBOVLN_SKIP=$sk python bench.py --steps 20 --warmup 5 --skip-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'], d['e2e']['ms_per_step'])"
done
