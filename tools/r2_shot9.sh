#!/bin/bash
# round 2, ninth GPU shot: P side with 1024 threads per row (with / without the columns kept in registers); transport at 4 CTAs/SM
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sweep.py -m gpu -x -q > gpurun_out/s9_sweep_tests.log 2>&1
tail -3 gpurun_out/s9_sweep_tests.log
rm -f gpurun_out/s9_sweep_bench.json
for cfg in "512 1" "1024 1" "1024 0" "512 0"; do
  set -- $cfg
  COGAPS_SWEEP_THREADS_LONG=$1 COGAPS_SWEEP_KEEP=$2 timeout 600 python tools/sweep_bench.py --ramp 500 --steps 20 >> gpurun_out/s9_sweep_bench.json 2>> gpurun_out/s9_sweep_bench.err
  echo "long=$1 keep=$2: $(tail -1 gpurun_out/s9_sweep_bench.json | cut -c1-420)"
done
