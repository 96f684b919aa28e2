#!/bin/bash
# round 2, twelfth GPU shot: rows handed out longest chain first (sweep_order_kernel)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep.py -m gpu -x -q > gpurun_out/s12_sweep_tests.log 2>&1; tail -2 gpurun_out/s12_sweep_tests.log
CFG="COGAPS_SWEEP_ORDER=0;;COGAPS_SWEEP_STAGE=2,COGAPS_SWEEP_THREADS=128,COGAPS_SWEEP_THREADS_LONG=512;COGAPS_SWEEP_STAGE=2,COGAPS_SWEEP_THREADS_LONG=512;COGAPS_SWEEP_SEG_FLOATS=5120;COGAPS_SWEEP_THREADS_LONG=512;COGAPS_SWEEP_ORDER=0;"
timeout 900 python tools/sweep_bench.py --ramp 500 --steps 20 --configs "$CFG" > gpurun_out/s12_sweep_bench.json 2> gpurun_out/s12_sweep_bench.err
cut -c1-330 gpurun_out/s12_sweep_bench.json
tail -5 gpurun_out/s12_sweep_bench.err
