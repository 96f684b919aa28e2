#!/bin/bash
# round 2, twenty-second GPU shot: bench.py as the driver runs it (N=1), both arms, final state
mkdir -p gpurun_out
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/s22_bench.json 2> gpurun_out/s22_bench.err
tail -c 300 gpurun_out/s22_bench.json; tail -3 gpurun_out/s22_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s22_bench_reference.json 2> gpurun_out/s22_bench_reference.err
tail -c 600 gpurun_out/s22_bench_reference.json
