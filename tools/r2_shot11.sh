#!/bin/bash
# round 2, eleventh GPU shot: what a row keeps in shared memory (D+AP / AP only) and clusters on the long rows
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep.py -m gpu -x -q > gpurun_out/s11_sweep_tests_default.log 2>&1; tail -2 gpurun_out/s11_sweep_tests_default.log
COGAPS_SWEEP_STAGE=2 timeout 900 python -m pytest tests/test_sweep.py -m gpu -x -q > gpurun_out/s11_sweep_tests_stage2.log 2>&1; tail -2 gpurun_out/s11_sweep_tests_stage2.log
COGAPS_SWEEP_SEG_FLOATS=5120 timeout 900 python -m pytest tests/test_sweep.py -m gpu -x -q > gpurun_out/s11_sweep_tests_cluster.log 2>&1; tail -2 gpurun_out/s11_sweep_tests_cluster.log
COGAPS_SWEEP_SEG_FLOATS=5120 COGAPS_SWEEP_CLUSTER=0 timeout 900 python -m pytest tests/test_sweep.py -m gpu -x -q > gpurun_out/s11_sweep_tests_segwalk.log 2>&1; tail -2 gpurun_out/s11_sweep_tests_segwalk.log
CFG=";COGAPS_SWEEP_STAGE=2;COGAPS_SWEEP_STAGE=2,COGAPS_SWEEP_THREADS=128,COGAPS_SWEEP_THREADS_LONG=512;COGAPS_SWEEP_STAGE=2,COGAPS_SWEEP_THREADS=128,COGAPS_SWEEP_THREADS_LONG=1024;COGAPS_SWEEP_THREADS=128;COGAPS_SWEEP_SEG_FLOATS=5120;COGAPS_SWEEP_SEG_FLOATS=6688;COGAPS_SWEEP_SEG_FLOATS=5120,COGAPS_SWEEP_THREADS_LONG=512;COGAPS_SWEEP_SEG_FLOATS=10240;"
timeout 900 python tools/sweep_bench.py --ramp 500 --steps 20 --configs "$CFG" > gpurun_out/s11_sweep_bench.json 2> gpurun_out/s11_sweep_bench.err
cut -c1-330 gpurun_out/s11_sweep_bench.json
tail -5 gpurun_out/s11_sweep_bench.err
