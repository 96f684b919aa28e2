#!/bin/bash
# round 2, thirty-first GPU shot: long rows read their D line past L1 (ld.global.cg)
mkdir -p gpurun_out
timeout 900 python tools/sweep_bench.py --ramp 500 --steps 20 --configs ";" > gpurun_out/s31_sweep_bench.json 2> gpurun_out/s31_sweep_bench.err
tail -2 gpurun_out/s31_sweep_bench.err
COGAPS_SWEEP_PROFILE=1 timeout 600 python tools/sweep_bench.py --ramp 500 --steps 2 > gpurun_out/s31_profile.json 2> gpurun_out/s31_profile.err
grep "sweep profile" gpurun_out/s31_profile.err | tail -2
timeout 900 python -m pytest tests/test_sweep.py -m gpu -x -q -k "long_rows or shared_memory or full_size" > gpurun_out/s31_tests.log 2>&1; tail -2 gpurun_out/s31_tests.log
