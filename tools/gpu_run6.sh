#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 900 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n 8 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run tests python -m pytest tests -m gpu -q --timeout 900 -x
run bench python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/bench_ops.json
run shapes python tools/ncu_shapes.py
run ncu_shapes ncu --profile-from-start off --set full --import-source on --clock-control none -k regex:gemm_tc -o gpurun_out/prof_shapes -f python tools/ncu_shapes.py
