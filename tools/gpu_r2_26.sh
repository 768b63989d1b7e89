#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/summary.txt
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -q -x --timeout 900 2>&1 | tail -3 | tee -a gpurun_out/summary.txt
run() { echo "=== $*" | tee -a gpurun_out/summary.txt; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline --profile-out gpurun_out/ops_$1.json 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])" | tee -a gpurun_out/summary.txt; }
run ROBOVLN_B_RESIDENT=1
run ROBOVLN_B_RESIDENT=0
timeout 900 python -m pytest tests -m gpu -q -x --timeout 900 2>&1 | tail -3 | tee -a gpurun_out/summary.txt
