#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/summary.txt
for mode in 1 5 9 13; do
  rm -f gpurun_out/gemm_times_m$mode.csv
  ROBOVLN_EPI_DEBUG=$mode ROBOVLN_GEMM_TIMES=gpurun_out/gemm_times_m$mode.csv timeout 600 python tools/gemm_timeline.py run > /dev/null 2>&1
  echo "=== mode $mode" | tee -a gpurun_out/summary.txt
  python tools/gemm_timeline.py show gpurun_out/gemm_times_m$mode.csv 2>&1 | grep "gemm\|T1\|c0\|c1\|all epi" | tee -a gpurun_out/summary.txt
done
