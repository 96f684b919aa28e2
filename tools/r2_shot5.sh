#!/bin/bash
# round 2, fifth GPU shot: the rest of the GPU suite (BASELINE-size / Tier-3 / adapter tests), smoke under ncu, full C3 run
mkdir -p gpurun_out
timeout 3000 python -m pytest tests -m gpu -q --durations=15 > gpurun_out/s5_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/s5_gpu_tests.log
tail -40 gpurun_out/s5_gpu_tests.log
t0=$(date +%s)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/s5_smoke_launches.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s5_smoke_ncu.log 2>&1
echo "smoke under ncu rc=$? seconds=$(( $(date +%s) - t0 ))" >> gpurun_out/s5_smoke_ncu.log
tail -4 gpurun_out/s5_smoke_ncu.log
cut -d, -f5 gpurun_out/s5_smoke_launches.csv | sort | uniq -c | sort -rn | head -24 > gpurun_out/s5_smoke_kernels.txt
timeout 1200 python tools/c3_full_run.py > gpurun_out/s5_c3_full_run.json 2> gpurun_out/s5_c3_full_run.err
echo "c3 full run rc=$?"
cat gpurun_out/s5_c3_full_run.json | cut -c1-1500
