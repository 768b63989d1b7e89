#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 600 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-400 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt gpurun_out/gemm_times_direct.csv
ROBOVLN_EPILOGUE=direct ROBOVLN_GEMM_TIMES=gpurun_out/gemm_times_direct.csv run tl_direct python tools/gemm_timeline.py run
run bench_tma python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline --profile-out gpurun_out/r2_11_ops_tma.json
ROBOVLN_EPILOGUE=direct run bench_direct python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline --profile-out gpurun_out/r2_11_ops_direct.json
ROBOVLN_EPILOGUE=direct TAILN=40 run probe_direct python tools/gemm_probe.py
