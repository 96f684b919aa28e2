#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/e2e_breakdown.py > gpurun_out/s23_e2e.log 2>&1
grep -v "host profile" gpurun_out/s23_e2e.log | tail -64
