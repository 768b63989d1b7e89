#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/epi
rm -f gpurun_out/epi/times_new.csv
ROBOVLN_GEMM_TIMES=gpurun_out/epi/times_new.csv timeout 300 python tools/gemm_timeline.py run > gpurun_out/epi/run_new.log 2>&1; echo rc=$?
timeout 600 python -m pytest tests/test_gemm_gpu.py -q -x --timeout 300 > gpurun_out/epi/t_gemm.log 2>&1; echo rc=$?; tail -3 gpurun_out/epi/t_gemm.log
