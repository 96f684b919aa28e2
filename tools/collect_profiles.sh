#!/bin/bash
# Round profile collection (run under gpurun on one GPU).  Kernel-level captures use the launch-per-batch
# mode (COGAPS_PERSISTENT=0): ncu serialises and replays launches, which a resident grid that converses with
# the host cannot survive; both modes run the same device code.
set -x
mkdir -p gpurun_out
python bench.py > gpurun_out/bench_full.json 2> gpurun_out/bench_full.err
tail -c 600 gpurun_out/bench_full.err
python bench.py --impl reference > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
COGAPS_PERSISTENT=0 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 400 --csv \
  --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 1 --ramp 30 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu.log 2>&1
COGAPS_PERSISTENT=0 ncu --set full --clock-control none --import-source on -k regex:eval_kernel -s 2500 -c 4 \
  -o gpurun_out/eval_kernel_full python bench.py --steps 1 --warmup 1 --ramp 30 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_full.log 2>&1
COGAPS_PERSISTENT=0 ncu --set full --clock-control none --import-source on -k regex:probe_kernel -c 4 \
  -o gpurun_out/probe_kernel_full python bench.py --steps 1 --warmup 1 --ramp 30 --no-e2e --no-cpu-baseline > gpurun_out/bench_under_ncu_probe.log 2>&1
ls -la gpurun_out | tail -14
