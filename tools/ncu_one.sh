#!/bin/bash
# usage: bash tools/ncu_one.sh <kernel regex> <skip> <count> <out tag>   (ncu --set full with sources on tests/dev_layer_times.py)
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o gpurun_out/$4 python tests/dev_layer_times.py > gpurun_out/$4.log 2>&1
ls -la gpurun_out/$4.ncu-rep
