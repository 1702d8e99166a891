#!/bin/bash
# source-level ncu captures (per-instruction counts and stall samples) of chosen conv launches of one training step
mkdir -p gpurun_out
for spec in "1 1 enc1b" "5 1 enc5" "21 2 head_dgrad"; do
  set -- $spec
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s $1 -c $2 -f -o gpurun_out/src2_$3 python tests/dev_layer_times.py > gpurun_out/src2_$3.log 2>&1
done
ls -la gpurun_out | grep src2
