#!/bin/bash
# round 2, nineteenth GPU shot: compute-sanitizer over every sampler / mode (incl. long rows and the row order kernel), smoke under
# ncu, the whole GPU suite, and configs[2] in full (10k + 10k iterations) on the final kernels
mkdir -p gpurun_out
for tool in memcheck racecheck synccheck initcheck; do
  timeout 1200 compute-sanitizer --tool $tool --print-limit 20 python tools/sanitize_small.py 8 > gpurun_out/s19_sanitize_$tool.log 2>&1
  echo "$tool rc=$? $(grep 'ERROR SUMMARY\|RACECHECK SUMMARY' gpurun_out/s19_sanitize_$tool.log | tail -1)"
done
t0=$(date +%s)
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/s19_smoke_launches.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s19_smoke_ncu.log 2>&1
echo "smoke under ncu rc=$? seconds=$(( $(date +%s) - t0 ))"
cut -d, -f5 gpurun_out/s19_smoke_launches.csv | sort | uniq -c | sort -rn | head -24 > gpurun_out/s19_smoke_kernels.txt
head -12 gpurun_out/s19_smoke_kernels.txt
timeout 2400 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/s19_gpu_tests.log 2>&1
echo "gpu tests rc=$?"; tail -14 gpurun_out/s19_gpu_tests.log
timeout 900 python tools/c3_full_run.py > gpurun_out/s19_c3_full_run.json 2> gpurun_out/s19_c3_full_run.err
cut -c1-400 gpurun_out/s19_c3_full_run.json
