#!/bin/bash
# round 2, twenty-ninth GPU shot: BASELINE configs[3] (sparse 50000 x 30000, 95 % zeros, k=50) through bench.py, both update modes
mkdir -p gpurun_out
timeout 1500 python bench.py --sparse --rows 50000 --cols 30000 --patterns 50 --steps 5 --warmup 3 --ramp 150 --no-e2e --no-cpu-baseline > gpurun_out/s29_sparse_c4.json 2> gpurun_out/s29_sparse_c4.err
echo "rc=$?"; tail -c 600 gpurun_out/s29_sparse_c4.json; tail -3 gpurun_out/s29_sparse_c4.err
