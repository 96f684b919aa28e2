#!/bin/bash
# round 2, twenty-seventh GPU shot: the whole GPU suite and bench.py (both arms) on the final state
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q --durations=6 > gpurun_out/s27_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -10 gpurun_out/s27_gpu_tests.log
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/s27_bench.json 2> gpurun_out/s27_bench.err
tail -c 300 gpurun_out/s27_bench.json; tail -3 gpurun_out/s27_bench.err
timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/s27_bench_reference.json 2> gpurun_out/s27_bench_reference.err
tail -c 400 gpurun_out/s27_bench_reference.json
