#!/bin/bash
# 2+ GPUs: the graphed training step with and without its one gradient all-reduce (SSDN_DEV_SKIP_ALLREDUCE=1 is a timing switch, not
# a valid training step), interleaved.  usage: bash tools/allreduce_cost.sh <gpus>
n=${1:-2}; port=29600
for r in 1 2; do
  for s in "" "SSDN_DEV_SKIP_ALLREDUCE=1"; do
    port=$((port+1))
    env $s timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $port bench.py --gpus $n --steps 30 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('%-28s %d GPUs: %.3f ms per step, %.0f patches/s' % ('${s:-with all-reduce}', d['n_gpus'], d['ms_per_step'], d['value']))"
  done
done
