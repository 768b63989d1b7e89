#!/bin/bash
# round 2, call 3: phase timeline of the fused block kernel; library arm; bf16 line; sanitizer evidence
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 900 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt gpurun_out/vla_times.csv
ROBOVLN_VLA_TIMES=gpurun_out/vla_times.csv run cfg3_times python bench.py --workload cross_modal --steps 2 --warmup 3
run lib python -m pytest tests/test_library_bar_gpu.py -q -x -k tuned -s
ROBOVLN_DTYPE=bf16 run bench_bf16 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline
run sanit_mem compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_vla_block_gpu.py -q -x -k "3-20 or 7-33"
run sanit_race compute-sanitizer --tool racecheck --error-exitcode 9 python -m pytest tests/test_vla_block_gpu.py -q -x -k "fp16 and 3-20"
