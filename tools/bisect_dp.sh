#!/bin/bash
# 2-GPU A/B of tests/dist_worker.py over several builds of the engine (ab_tmp/*.so, untracked)
port=29560
for lib in "$@"; do
  port=$((port+1))
  echo "== $lib"
  SSDN_LIB=$lib python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $port tests/dist_worker.py 2>&1 | grep MULTIRANK | python -c "
import sys, json
for l in sys.stdin:
    d = json.loads(l[len('MULTIRANK '):])
    for k, v in d['cases'].items(): print(' ', k, 'weights_rel_l2 %.2e' % v['weights_rel_l2'], v['weights_worst'].split('.')[-2:], 'grad %.2e' % v['first_gradient_rel_l2'], 'loss %.1e' % v['loss_rel_diff'])
"
done
