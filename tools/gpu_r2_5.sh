#!/bin/bash
# round 2, call 5 (2 GPUs): two-device module placement, strong-scaling bench at N=2, reference arm, new tests
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run() { local name=$1; shift; echo "=== $name: $*" | tee -a gpurun_out/summary.txt; timeout 1200 "$@" > gpurun_out/$name.log 2>&1; local rc=$?; echo "rc=$rc" | tee -a gpurun_out/summary.txt; tail -n ${TAILN:-6} gpurun_out/$name.log | tee -a gpurun_out/summary.txt; }
rm -f gpurun_out/summary.txt
run t_two python -m pytest tests/test_parity_gpu.py -q --timeout 600 -x -k "two_devices or fp16_range or bf16_build or instruction_cache"
run bench2 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5
run bench2w python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --gpus 2 --steps 20 --warmup 5 --scaling weak
run ref1 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1
run bench1 python bench.py --steps 20 --warmup 5
