#!/bin/bash
cd "$(dirname "$0")/.."
O=gpurun_out/final; mkdir -p $O
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > $O/clocks.csv &
SMI=$!
timeout 1200 python bench.py --steps 20 --warmup 5 --profile-out $O/r02_bench_ops.json > $O/bench.log 2>&1
kill $SMI
grep "^{" $O/bench.log | tail -1 | cut -c1-300
timeout 600 python bench.py > $O/bench_default.log 2>&1
grep "^{" $O/bench_default.log | tail -1 | cut -c1-300
timeout 900 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
