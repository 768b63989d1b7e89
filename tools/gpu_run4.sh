#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 600 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n 12 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run pair python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 300 -x -k "pair"
run tests python -m pytest tests -m gpu -q --timeout 900 -x
run bench python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/bench_ops.json
ROBOVLN_PAIR_MMA=0 run bench_nopair python bench.py --steps 20 --warmup 5 --skip-cpu-baseline
