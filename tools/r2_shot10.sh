#!/bin/bash
# round 2, tenth GPU shot: compute-sanitizer over small runs of every sampler / mode
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py 8 > gpurun_out/s10_sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep -c 'ERROR SUMMARY' gpurun_out/s10_sanitize_$tool.log) $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/s10_sanitize_$tool.log | tail -2)"
done
timeout 600 python -m pytest tests/test_sweep.py -m gpu -x -q 2>&1 | tail -2
timeout 300 python tools/sweep_bench.py --ramp 500 --steps 20 | cut -c1-400
