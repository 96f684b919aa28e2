#!/bin/bash
# round 2, twenty-first GPU shot: cgb_run's setup with lambda's running sums taken off the chain of fp adds (accumulateRun) and the
# strips gathered ahead by helper threads; a few chain tests to see lambda's bits did not move
mkdir -p gpurun_out
timeout 600 python tools/e2e_breakdown.py > gpurun_out/s21_e2e.log 2>&1
grep -v "host profile" gpurun_out/s21_e2e.log | tail -12
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "run_matches_oracle or posterior_means" > gpurun_out/s21_tests.log 2>&1; tail -2 gpurun_out/s21_tests.log
