#!/bin/bash
# round 2, call 8: pair kernel with pipelined LayerNorm epilogues, stream-parallel stand-alone stage
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 600 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-8} gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt gpurun_out/vla_times.csv
run t_vla python -m pytest tests/test_vla_block_gpu.py tests/test_train_kernels_gpu.py -q --timeout 300 -x
run cfg3_pair python bench.py --workload cross_modal --steps 100 --warmup 10
ROBOVLN_VLA_PAIR=0 run cfg3_single python bench.py --workload cross_modal --steps 100 --warmup 10
ROBOVLN_VLA_TIMES=gpurun_out/vla_times.csv run cfg3_times python bench.py --workload cross_modal --steps 2 --warmup 3
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__grid_size,launch__registers_per_thread
run ncu_cfg3 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_cfg3_ncu.csv python bench.py --workload cross_modal --steps 2 --warmup 3
run t_parity python -m pytest tests/test_parity_gpu.py -q --timeout 600 -x
run sanit_pair compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_vla_block_gpu.py -q -x -k "pair and (3-20 or 7-33)"
