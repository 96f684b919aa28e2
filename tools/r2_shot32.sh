#!/bin/bash
# round 2, last GPU shot: the whole GPU suite at HEAD
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/s32_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -4 gpurun_out/s32_gpu_tests.log
