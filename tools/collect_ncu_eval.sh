#!/bin/bash
# the two ncu passes over the per-batch eval kernel (launch-per-batch mode, chain ramped for 30 iterations so the
# batches have their usual 100-300 tasks)
set -x
COGAPS_PERSISTENT=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --ramp 30 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
COGAPS_PERSISTENT=0 ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 2500 -c 4 \
  -o gpurun_out/eval_kernel_full python bench.py --steps 1 --warmup 1 --ramp 30 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
ls -la gpurun_out | grep -i "ncu-rep\|launches"
