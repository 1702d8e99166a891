#!/bin/bash
# usage: bash tools/sweep_bench.sh <rounds> "VAR=a" "VAR=b" ...   ("none" = defaults)
# interleaved bench.py runs (the chip's clocks drift with temperature / the power cap): prints resident and end-to-end patches/s
rounds=$1; shift
mkdir -p gpurun_out
for r in $(seq 1 $rounds); do
  for setting in "$@"; do
    if [ "$setting" == "none" ]; then s=""; else s="$setting"; fi
    env $s python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-extras 2>/dev/null | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.read())
print('%-40s resident %8.1f  e2e %8.1f patches/s   %.3f ms' % ('$setting', d['value'], d['e2e']['value'], d['ms_per_step']))"
  done
done
