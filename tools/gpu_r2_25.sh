#!/bin/bash
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
rm -f gpurun_out/summary.txt
run() { echo "=== $*" | tee -a gpurun_out/summary.txt; env "$@" timeout 600 python bench.py --steps 20 --warmup 5 --skip-cpu-baseline --skip-strong-base --skip-library-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['ms_per_step'])" | tee -a gpurun_out/summary.txt; }
run A=1
run ROBOVLN_LN_FUSED=2
run ROBOVLN_LN_FUSED=3
run ROBOVLN_ATTN=tc
run ROBOVLN_PAIR_MMA=0
run ROBOVLN_PDL=1
run ROBOVLN_RGB_SPLIT=2
run ROBOVLN_GN_EPILOGUE=0
run ROBOVLN_PRIORITIES=0
