#!/bin/bash
# round 2, call 1: new fused kernel + prep kernels + bench contract
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 900 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n 12 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run t_vla python -m pytest tests/test_vla_block_gpu.py -q --timeout 600 -x
run t_all python -m pytest tests -m gpu -q --timeout 900 -x --deselect tests/test_vla_block_gpu.py
run smoke python -c "import __graft_entry__ as g; g.smoke()"
run bench python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/r2_1_ops.json
run cfg3 python bench.py --workload cross_modal --steps 50 --warmup 5
