#!/bin/bash
# round 2, call 2: fused block kernel with rotated weight walk / 5-slot ring; GEMM k-rotation probe; ncu of cfg3
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 900 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run t_vla python -m pytest tests/test_vla_block_gpu.py tests/test_parity_gpu.py -q --timeout 600 -x
run cfg3_rot1 python bench.py --workload cross_modal --steps 100 --warmup 10
ROBOVLN_VLA_ROTATE=0 run cfg3_rot0 python bench.py --workload cross_modal --steps 100 --warmup 10
run bench_krot0 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --profile-out gpurun_out/r2_2_ops_krot0.json
ROBOVLN_GEMM_KROT=1 run bench_krot1 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --profile-out gpurun_out/r2_2_ops_krot1.json
TAILN=40 run probe_krot0 python tools/gemm_probe.py
ROBOVLN_GEMM_KROT=1 TAILN=40 run probe_krot1 python tools/gemm_probe.py
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,lts__t_sectors_srcunit_tex_op_read.sum,lts__throughput.avg.pct_of_peak_sustained_elapsed,sm__throughput.avg.pct_of_peak_sustained_elapsed,launch__grid_size,launch__registers_per_thread,smsp__cycles_active.avg
run ncu_cfg3 ncu --metrics $M --clock-control none --csv --log-file gpurun_out/r2_cfg3_ncu.csv python bench.py --workload cross_modal --steps 2 --warmup 3
run ncu_cfg3_full ncu --set full --import-source on --clock-control none -k regex:vla_block -s 3 -c 2 -o gpurun_out/r2_vla_full -f python bench.py --workload cross_modal --steps 2 --warmup 3
