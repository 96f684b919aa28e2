#!/bin/bash
# round 2, fourteenth GPU shot: ncu --set full of the sweep kernel (A side 256 threads / P side 512 threads, AP line only) at steady state
mkdir -p gpurun_out
timeout 1500 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 1002 -c 2 -f -o gpurun_out/s14_sweep_full \
  python tools/sweep_bench.py --ramp 500 --steps 2 > gpurun_out/s14_ncu.log 2>&1
tail -5 gpurun_out/s14_ncu.log
ls -la gpurun_out/s14_sweep_full.ncu-rep
