#!/bin/bash
# round 2, twenty-eighth GPU shot: ncu --set full of the whole-matrix passes (transpose both ways, chi-square, AP rebuild) at configs[2]'s size
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"transpose_kernel|chisq_kernel|rebuild_ap_kernel" -s 70 -c 6 -f -o gpurun_out/s28_passes \
  python tools/sync_time.py > gpurun_out/s28_ncu.log 2>&1
tail -4 gpurun_out/s28_ncu.log
