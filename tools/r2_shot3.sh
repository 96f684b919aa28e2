#!/bin/bash
# round 2, third GPU shot: sweep tests (70000-long rows incl.), smoke under ncu (launch list), bench N=1
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep.py -m gpu -x -q > gpurun_out/s3_sweep_tests.log 2>&1
echo "sweep tests rc=$?" >> gpurun_out/s3_sweep_tests.log
tail -4 gpurun_out/s3_sweep_tests.log
ncu python -c 'import os;print({k:v for k,v in os.environ.items() if "INJ" in k or "NSIGHT" in k or "COMPUTE" in k or "PROF" in k or "LD_PRELOAD" in k})' > gpurun_out/s3_ncu_env.log 2>&1
t0=$(date +%s)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/s3_smoke_launches.csv \
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke_ncu.log 2>&1
echo "smoke under ncu rc=$? seconds=$(( $(date +%s) - t0 ))" >> gpurun_out/s3_smoke_ncu.log
tail -5 gpurun_out/s3_smoke_ncu.log
cut -d, -f5 gpurun_out/s3_smoke_launches.csv | sort | uniq -c | sort -rn | head -20 > gpurun_out/s3_smoke_kernels.txt
cat gpurun_out/s3_smoke_kernels.txt
CUDA_LAUNCH_BLOCKING=1 timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s3_smoke_blocking.log 2>&1
echo "smoke CUDA_LAUNCH_BLOCKING rc=$?" >> gpurun_out/s3_smoke_blocking.log
tail -3 gpurun_out/s3_smoke_blocking.log
timeout 1500 python bench.py --steps 20 --warmup 3 > gpurun_out/s3_bench.json 2> gpurun_out/s3_bench.err
echo "bench rc=$?"
tail -c 3000 gpurun_out/s3_bench.json
tail -5 gpurun_out/s3_bench.err
