#!/bin/bash
# One GPU-box visit: parity tests, bench, per-layer times, ncu launch list and ncu full captures.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [skip_tests]
TAG=${1:-r01}
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/${TAG}_build.log 2>&1
if [ -z "$2" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -3 gpurun_out/${TAG}_pytest.log
fi
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
SMI=$!
timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
kill $SMI
cat gpurun_out/${TAG}_bench.json
timeout 300 python tests/dev_layer_times.py > gpurun_out/${TAG}_layer_times.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${TAG}_launches.csv \
   python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/${TAG}_ncu_bench.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 15 -c 2 -f -o gpurun_out/${TAG}_conv_fwd \
   python tests/dev_layer_times.py > gpurun_out/${TAG}_ncu_conv.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_igemm_kernel -s 3 -c 1 -f -o gpurun_out/${TAG}_wgrad \
   python tests/dev_layer_times.py > gpurun_out/${TAG}_ncu_wgrad.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:wgrad_small_cin -c 1 -f -o gpurun_out/${TAG}_smallcin \
   python tests/dev_layer_times.py > gpurun_out/${TAG}_ncu_smallcin.log 2>&1
SSDN_CONV_STATS=1 timeout 200 python tests/dev_layer_times.py > gpurun_out/${TAG}_role_waits.log 2>&1
ls -la gpurun_out
