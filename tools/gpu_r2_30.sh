#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout 600 python tools/gemm_probe.py > gpurun_out/probe_b2b.log 2>&1
grep "act=0\|cublas" gpurun_out/probe_b2b.log | grep -v "bn128"
for sk in none depth rgb bert; do
  ROBOVLN_SKIP=$sk timeout 300 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('skip=$sk', round(d['ms_per_step'],3))"
done
