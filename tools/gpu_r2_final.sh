#!/bin/bash
# round 2: the evidence set copied into profiles/ (1 GPU)
cd "$(dirname "$0")/.."
O=gpurun_out/final; mkdir -p $O
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a $O/summary.txt; timeout 1200 "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/summary.txt; tail -n 2 $O/$name.log | cut -c1-300 | tee -a $O/summary.txt; }
rm -f $O/summary.txt $O/vla_times.csv
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/clocks.csv &
SMI=$!
run bench python bench.py --steps 20 --warmup 5 --profile-out $O/r02_bench_ops.json
kill $SMI
run reference python bench.py --impl reference --steps 5 --warmup 1
ROBOVLN_DTYPE=bf16 run bench_bf16 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline
run cfg3 python bench.py --workload cross_modal --steps 100 --warmup 10
ROBOVLN_VLA_PAIR=0 run cfg3_single python bench.py --workload cross_modal --steps 100 --warmup 10
run cfg5 python bench.py --workload train --steps 20 --warmup 5
run icache python tools/instr_cache_probe.py
run probe python tools/gemm_probe.py
ROBOVLN_VLA_TIMES=$O/vla_times.csv run cfg3_times python bench.py --workload cross_modal --steps 2 --warmup 3
M=gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,launch__grid_size,launch__registers_per_thread
run ncu_cfg3 ncu --metrics $M --clock-control none --csv --log-file $O/r02_cfg3_ncu.csv python bench.py --workload cross_modal --steps 2 --warmup 3
run ncu_cfg3_full ncu --set full --import-source on --clock-control none -k regex:vla_pair -s 3 -c 1 -o $O/r02_vla_pair_full -f python bench.py --workload cross_modal --steps 2 --warmup 3
M2=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ROBOVLN_MULTISTREAM=0 run ncu_step ncu --profile-from-start off --metrics $M2 --clock-control none --cache-control none --csv --log-file $O/r02_ncu_launches.csv python tools/ncu_step.py
run sanit_mem compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_vla_block_gpu.py tests/test_train_kernels_gpu.py tests/test_kernels_gpu.py -q -x -k "3-20 or 7-33 or loss or adam or lstm or groupnorm or layernorm"
run sanit_mem_policy compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_parity_gpu.py -q -x -k "cfg1_b2 or graph_replay or instruction_cache"
run sanit_race compute-sanitizer --tool racecheck python -m pytest tests/test_vla_block_gpu.py -q -x -k "fp16 and 3-20"
