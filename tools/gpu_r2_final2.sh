#!/bin/bash
# round 2, last evidence pass after the final kernel changes (1 GPU): contract line (default flags), 20-step line + per-launch
# profile, bf16 line, cfg3, cfg5, instruction-cache probe, reference arm, ncu launch list of one step, full GPU test suite
cd "$(dirname "$0")/.."
O=gpurun_out/final; mkdir -p $O
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a $O/summary.txt; timeout 1200 "$@" > $O/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a $O/summary.txt; tail -n 2 $O/$name.log | cut -c1-300 | tee -a $O/summary.txt; }
rm -f $O/summary.txt
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/clocks.csv &
SMI=$!
run bench python bench.py --steps 20 --warmup 5 --profile-out $O/r02_bench_ops.json
kill $SMI
run bench_default python bench.py
run reference python bench.py --impl reference --steps 5 --warmup 1
ROBOVLN_DTYPE=bf16 run bench_bf16 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline
run cfg3 python bench.py --workload cross_modal --steps 100 --warmup 10
ROBOVLN_VLA_PAIR=0 run cfg3_single python bench.py --workload cross_modal --steps 100 --warmup 10
run cfg5 python bench.py --workload train --steps 20 --warmup 5
run icache python tools/instr_cache_probe.py
M2=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active
ROBOVLN_MULTISTREAM=0 run ncu_step ncu --profile-from-start off --metrics $M2 --clock-control none --cache-control none --csv --log-file $O/r02_ncu_launches.csv python tools/ncu_step.py
run tests python -m pytest tests -m gpu -q --timeout 900
