#!/bin/bash
# round 2, twentieth GPU shot: where cgb_run's setup goes on the GPU box; lambda's column-walking running sum with / without helper threads
mkdir -p gpurun_out
COGAPS_SUM_PIPELINE=0 timeout 600 python tools/e2e_breakdown.py > gpurun_out/s20_e2e_plain.log 2>&1
grep -v "host profile" gpurun_out/s20_e2e_plain.log | tail -9
timeout 600 python tools/e2e_breakdown.py > gpurun_out/s20_e2e_pipelined.log 2>&1
grep -v "host profile" gpurun_out/s20_e2e_pipelined.log | tail -9
