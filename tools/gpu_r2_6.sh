#!/bin/bash
# round 2, call 6: training-tail kernels, cfg5 line
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run t_train python -m pytest tests/test_train_kernels_gpu.py tests/test_training_tail.py -q --timeout 600 -x
run cfg5 python bench.py --workload train --steps 20 --warmup 5
