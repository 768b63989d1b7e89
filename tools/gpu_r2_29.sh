#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/summary.txt
timeout 600 python -m pytest tests/test_vla_block_gpu.py -q -x --timeout 600 2>&1 | tail -2 | tee -a gpurun_out/summary.txt
for v in 1 0; do
ROBOVLN_VLA_PAIR=$v timeout 300 python bench.py --workload cross_modal --steps 100 --warmup 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg3 pair=$v', d['ms_per_step'])" | tee -a gpurun_out/summary.txt
done
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('step', d['ms_per_step'], 'e2e', d['e2e']['ms_per_step'])" | tee -a gpurun_out/summary.txt
timeout 900 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -2 | tee -a gpurun_out/summary.txt
