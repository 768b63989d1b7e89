#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 900 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run t_gemm python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -q --timeout 600 -x
run t_parity python -m pytest tests/test_parity_gpu.py -q --timeout 600 -x
run bench python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline --profile-out gpurun_out/r2_14_ops.json
TAILN=40 run probe python tools/gemm_probe.py
