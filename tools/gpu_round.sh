#!/bin/bash
# One GPU-box visit: clean build, parity tests, bench, per-layer times, ncu launch list and ncu full captures, sanitizers.
# usage (from the repo root, under gpurun): bash tools/gpu_round.sh <tag> [what...]     what: build test bench layers ncu sanitize (default: all)
TAG=${1:-r02}
shift
WHAT=${*:-build test bench layers ncu sanitize}
mkdir -p gpurun_out
has() { [[ " $WHAT " == *" $1 "* ]]; }
if has build; then
  # a CLEAN nvcc build on the GPU box (the snapshot ships a prebuilt .so; this proves the sources build there too)
  python -c "import __graft_entry__ as g; g.build(force=True)" > gpurun_out/${TAG}_build.log 2>&1; echo "build exit $?" >> gpurun_out/${TAG}_build.log
  tail -2 gpurun_out/${TAG}_build.log
fi
if has test; then
  timeout 1200 python -m pytest tests -m gpu -q --durations=8 > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/${TAG}_pytest.log
  tail -4 gpurun_out/${TAG}_pytest.log
fi
if has bench; then
  nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap --format=csv -lms 200 > gpurun_out/${TAG}_clocks.csv &
  SMI=$!
  timeout 600 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench exit $?"
  kill $SMI
  python - <<PY
import json
d = json.loads([l for l in open("gpurun_out/${TAG}_bench.json") if l.startswith("{")][-1])
print({k: d[k] for k in ("value", "ms_per_step", "e2e", "stale_scale_passes", "cpu_baseline")})
r = d["roofline"]; print({k: r[k] for k in ("kernel", "achieved", "peak", "frac", "issued_frac", "traffic")})
for k, v in r["kernels"].items(): print(" ", k, {a: (round(b, 3) if isinstance(b, float) else b) for a, b in v.items()})
for k, v in d.get("configs", {}).items(): print(" ", k, round(v["value"]), "patches/s", round(v["ms_per_step"], 3), "ms")
PY
fi
if has layers; then
  timeout 300 python tests/dev_layer_times.py > gpurun_out/${TAG}_layer_times.log 2>&1
  SSDN_CONV_STATS=1 timeout 200 python tests/dev_layer_times.py > gpurun_out/${TAG}_role_waits.log 2>&1
fi
if has ncu; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/${TAG}_launches.csv \
     python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-extras --no-graph > gpurun_out/${TAG}_ncu_bench.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:conv_igemm_kernel -s 15 -c 2 -f -o gpurun_out/${TAG}_conv_fwd \
     python tests/dev_layer_times.py > gpurun_out/${TAG}_ncu_conv.log 2>&1
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:wgrad_igemm_kernel -s 3 -c 2 -f -o gpurun_out/${TAG}_wgrad \
     python tests/dev_layer_times.py > gpurun_out/${TAG}_ncu_wgrad.log 2>&1
  timeout 600 ncu --set full --clock-control none -k regex:"pool_bwd|up_bwd|pool_fwd|wgrad_reduce_batched|pack_nchw_pixel|pack_input|adam|posterior" -c 16 -f -o gpurun_out/${TAG}_pointwise \
     python tests/dev_layer_times.py > gpurun_out/${TAG}_ncu_pointwise.log 2>&1
fi
if has sanitize; then
  # one small training step (smoke(): 2 x 3 x 32 x 32, blind-spot net + posterior + Adam, ~110 launches) under each tool
  for tool in memcheck synccheck racecheck; do
    timeout 900 compute-sanitizer --tool $tool --print-limit 20 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${TAG}_sanitizer_${tool}.log 2>&1
    echo "$tool exit $?" >> gpurun_out/${TAG}_sanitizer_${tool}.log
    grep -E "ERROR SUMMARY|exit|smoke\]" gpurun_out/${TAG}_sanitizer_${tool}.log | tail -3
  done
fi
ls gpurun_out | grep ${TAG}_ | tr '\n' ' '
