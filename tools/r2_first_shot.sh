#!/bin/bash
# First gpurun call of a round (one GPU): the whole GPU test suite, the two bench arms, and where the end-to-end call
# spends its time.  Everything lands in gpurun_out/.
#   gpurun --timeout 1500 -- 'bash tools/r2_first_shot.sh'
set -x
mkdir -p gpurun_out
python -m pytest tests -q -m gpu -p no:cacheprovider > gpurun_out/gpu_tests.log 2>&1; tail -3 gpurun_out/gpu_tests.log
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err; tail -c 400 gpurun_out/bench_full.err
COGAPS_HOST_PROFILE=1 python tools/e2e_breakdown.py > gpurun_out/e2e_breakdown.log 2>&1; tail -8 gpurun_out/e2e_breakdown.log
ls -la gpurun_out | tail -8
