#!/bin/bash
# round 2: strong scaling at 8, 4 and 2 GPUs (fixed 512-environment global batch), run on one 8-GPU box
cd "$(dirname "$0")/.."
O=gpurun_out/scale; mkdir -p $O
export PYTHONUNBUFFERED=1
for N in 8 4 2; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 > $O/bench$N.log 2>&1
  echo "N=$N rc=$?"; grep "^{" $O/bench$N.log | tail -1 | cut -c1-200
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29650 bench.py --gpus 2 --steps 20 --warmup 5 --scaling weak > $O/bench2_weak.log 2>&1
echo "N=2 weak rc=$?"; grep "^{" $O/bench2_weak.log | tail -1 | cut -c1-200
timeout 600 python -m pytest tests/test_parity_gpu.py -q -x -k "two_devices" 2>&1 | tail -2
