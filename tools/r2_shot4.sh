#!/bin/bash
# round 2, fourth GPU shot: whole GPU suite with the BASELINE-size / Tier-3 tests, smoke under ncu, e2e breakdown
mkdir -p gpurun_out
t0=$(date +%s)
timeout 1500 ncu --metrics gpu__time_duration.sum --clock-control none -c 20000 --csv --log-file gpurun_out/s4_smoke_launches.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s4_smoke_ncu.log 2>&1
echo "smoke under ncu rc=$? seconds=$(( $(date +%s) - t0 ))" >> gpurun_out/s4_smoke_ncu.log
tail -5 gpurun_out/s4_smoke_ncu.log
cut -d, -f5 gpurun_out/s4_smoke_launches.csv | sort | uniq -c | sort -rn | head -24 > gpurun_out/s4_smoke_kernels.txt
cat gpurun_out/s4_smoke_kernels.txt
timeout 2400 python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/s4_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/s4_gpu_tests.log
tail -25 gpurun_out/s4_gpu_tests.log
COGAPS_HOST_PROFILE=1 timeout 600 python - > gpurun_out/s4_e2e_breakdown.log 2>&1 <<'PY'
import time, sys
sys.path.insert(0, '.')
import bench, cogaps_b200 as cg
data = bench.make_data()
for mode in (1, 1, 0):
    t0 = time.perf_counter()
    res = cg.gaps_run(data, seed=42, nPatterns=20, nIterations=100, outputFrequency=0, maxThreads=1, updateMode=mode)
    print("mode", mode, "wall %.3f s" % (time.perf_counter() - t0), "loop %.3f" % res.totalRunningTime, "updates", res.totalUpdates, flush=True)
PY
cat gpurun_out/s4_e2e_breakdown.log | grep -v "host profile" | tail -12
