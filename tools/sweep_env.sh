#!/bin/bash
# usage: bash tools/sweep_env.sh "VAR=a VAR=b ..."   (each token is one environment setting; "none" = defaults)
# runs tests/dev_layer_times.py under each and prints the per-kind totals + the big layers
mkdir -p gpurun_out
for setting in "$@"; do
  tag=$(echo "$setting" | tr ' =/' '___')
  if [ "$setting" == "none" ]; then env python tests/dev_layer_times.py > gpurun_out/sweep_$tag.log 2>&1
  else env $setting python tests/dev_layer_times.py > gpurun_out/sweep_$tag.log 2>&1; fi
  echo "== $setting"
  grep -E "^conv_fwd +(enc1a|enc1b|dec2b|dec1a|dec1b|head1|head2)" gpurun_out/sweep_$tag.log | awk '{printf "%s %s  ", $2, $3}'; echo
  grep -A20 "per kind" gpurun_out/sweep_$tag.log | grep -E "conv_fwd|conv_dgrad|wgrad |total" | awk '{printf "%s %s %s | ", $1, $4, $5}'; echo
done
