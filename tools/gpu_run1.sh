#!/bin/bash
# First GPU session: unit tests per group (separate processes so a trapped kernel does not poison
# the rest), end-to-end parity with the validation GEMM and with tcgen05, smoke, bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw,memory.total --format=csv > gpurun_out/smi.txt 2>&1
run() { # name, cmd...
  local name=$1; shift
  echo "=== $name: $*" | tee -a gpurun_out/summary.txt
  timeout 900 "$@" > gpurun_out/$name.log 2>&1
  local rc=$?
  echo "rc=$rc" | tee -a gpurun_out/summary.txt
  tail -n 6 gpurun_out/$name.log | tee -a gpurun_out/summary.txt
}
run build python -c "import __graft_entry__ as G; G.build(); print('build ok')"
run kernels python -m pytest tests/test_kernels_gpu.py -m gpu -q --timeout 600
run gemm_plain python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 600 -k "plain_tail_m300 or plain_layer1"
run gemm_conv python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 600 -k "c3x3_64x64_c64 or c3x3_32x32_c128"
run gemm_stride python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 600 -k "c3x3_s2_64_to_32 or c1x1_s2_ds_64"
run gemm_all python -m pytest tests/test_gemm_gpu.py -m gpu -q --timeout 900
ROBOVLN_GEMM=simt run parity_simt python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 900
ROBOVLN_MULTISTREAM=0 run parity_tc_1stream python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 900
run parity_tc python -m pytest tests/test_parity_gpu.py -m gpu -q --timeout 900
run smoke python -c "import __graft_entry__ as G; G.smoke()"
run bench python bench.py --steps 10 --warmup 3 --profile-out gpurun_out/bench_ops.json
cat gpurun_out/summary.txt
