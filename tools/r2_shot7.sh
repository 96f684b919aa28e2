#!/bin/bash
# round 2, seventh GPU shot: sweep with the decision inlined; 128 vs 256 threads per row on the A side; sweep tests
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sweep.py -m gpu -x -q > gpurun_out/s7_sweep_tests.log 2>&1
tail -3 gpurun_out/s7_sweep_tests.log
for T in 0 128 256; do
  COGAPS_SWEEP_THREADS=$T timeout 600 python tools/sweep_bench.py --ramp 500 --steps 20 >> gpurun_out/s7_sweep_bench.json 2>> gpurun_out/s7_sweep_bench.err
  tail -1 gpurun_out/s7_sweep_bench.json | cut -c1-700
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1000 -c 2 -o gpurun_out/s7_sweep_full -f \
  python tools/sweep_bench.py --ramp 500 --steps 3 > gpurun_out/s7_ncu_full.log 2>&1
tail -2 gpurun_out/s7_ncu_full.log
