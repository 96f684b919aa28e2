#!/bin/bash
# round 2, twenty-sixth GPU shot: the AP transpose with 64 x 64 tiles and 16-byte accesses
mkdir -p gpurun_out
timeout 600 python tools/sync_time.py > gpurun_out/s26_sync.log 2>&1; tail -4 gpurun_out/s26_sync.log
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_sweep.py -m gpu -x -q -k "full_size or run_matches_oracle or long_rows or shared_memory" > gpurun_out/s26_tests.log 2>&1; tail -2 gpurun_out/s26_tests.log
