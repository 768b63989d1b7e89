#!/bin/bash
# round 2: strong scaling at 8 and 4 GPUs (fixed 512-environment global batch), run on one 8-GPU box
cd "$(dirname "$0")/.."
O=gpurun_out/scale; mkdir -p $O
export PYTHONUNBUFFERED=1
for N in 8 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $((29600+N)) bench.py --gpus $N --steps 20 --warmup 5 > $O/bench$N.log 2>&1
  echo "N=$N rc=$?"; grep "^{" $O/bench$N.log | tail -1 | cut -c1-200
done
