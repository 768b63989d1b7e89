#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/epi
rm -f gpurun_out/epi/times_new2.csv
ROBOVLN_GEMM_TIMES=gpurun_out/epi/times_new2.csv timeout 300 python tools/gemm_timeline.py run > gpurun_out/epi/run_new2.log 2>&1; echo rc=$?
timeout 600 python -m pytest tests/test_gemm_gpu.py tests/test_kernels_gpu.py -q -x --timeout 300 > gpurun_out/epi/t_gemm.log 2>&1; echo rc=$?; tail -3 gpurun_out/epi/t_gemm.log
timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline --profile-out gpurun_out/epi/ops_stamps.json > gpurun_out/epi/bench_stamps.log 2>&1; tail -1 gpurun_out/epi/bench_stamps.log | cut -c1-200
