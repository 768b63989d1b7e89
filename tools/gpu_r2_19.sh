#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 900 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-4} gpurun_out/$name.log | cut -c1-300 | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run bench python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline --profile-out gpurun_out/r2_19_ops.json
run bench_b python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline
run t_all python -m pytest tests -m gpu -q --timeout 900 -x
