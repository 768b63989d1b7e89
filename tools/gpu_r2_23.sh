#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/summary.txt gpurun_out/gemm_times_conv.csv
ROBOVLN_GEMM_TIMES=gpurun_out/gemm_times_conv.csv timeout 600 python tools/gemm_timeline.py run conv > gpurun_out/tl.log 2>&1
tail -3 gpurun_out/tl.log
python tools/gemm_timeline.py show gpurun_out/gemm_times_conv.csv 2>&1 | tee -a gpurun_out/summary.txt
