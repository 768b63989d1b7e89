#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 900 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n 8 gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run gemm python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 900 -x
ROBOVLN_EPILOGUE=direct run gemm_direct python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 900
run tests python -m pytest tests -m gpu -q --timeout 900
run smoke python -c "import __graft_entry__ as G; G.smoke()"
run bench python bench.py --steps 20 --warmup 5 --profile-out gpurun_out/bench_ops.json
