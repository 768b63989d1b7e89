#!/bin/bash
# which part of the gemm_tc epilogue costs the time?  (ROBOVLN_EPI_DEBUG bits: 1 no store, 2 no math, 4 no TMEM load) -- stamps build
cd "$(dirname "$0")/.."
mkdir -p gpurun_out/epi
for d in 0 1 2 4 7; do
  rm -f gpurun_out/epi/times_$d.csv
  ROBOVLN_EPI_DEBUG=$d ROBOVLN_GEMM_TIMES=gpurun_out/epi/times_$d.csv timeout 300 python tools/gemm_timeline.py run > gpurun_out/epi/run_$d.log 2>&1
  echo "dbg=$d rc=$?"
done
