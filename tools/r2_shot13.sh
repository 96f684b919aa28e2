#!/bin/bash
# round 2, thirteenth GPU shot: the whole GPU suite on the new defaults (long rows: AP line only, 512 threads; rows longest chain
# first), the phase profile of a row's CTA, and the sweep timing
mkdir -p gpurun_out
timeout 1800 python -m pytest tests -m gpu -x -q > gpurun_out/s13_gpu_tests.log 2>&1; tail -3 gpurun_out/s13_gpu_tests.log
CFG=";COGAPS_SWEEP_STAGE=1,COGAPS_SWEEP_THREADS_LONG=1024;COGAPS_SWEEP_ORDER=0;"
timeout 900 python tools/sweep_bench.py --ramp 500 --steps 20 --configs "$CFG" > gpurun_out/s13_sweep_bench.json 2> gpurun_out/s13_sweep_bench.err
tail -5 gpurun_out/s13_sweep_bench.err
COGAPS_SWEEP_PROFILE=1 timeout 600 python tools/sweep_bench.py --ramp 500 --steps 2 > gpurun_out/s13_profile.json 2> gpurun_out/s13_profile.err
grep "sweep profile" gpurun_out/s13_profile.err | tail -2
