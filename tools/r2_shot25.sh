#!/bin/bash
# round 2, twenty-fifth GPU shot (8 GPUs): bench.py --gpus 8 as the driver launches it, final state (identical replicas, C5 record
# with the gather of the row copy, per-rank times)
mkdir -p gpurun_out
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/s25_bench_8gpu.json 2> gpurun_out/s25_bench_8gpu.err
echo "bench --gpus 8 rc=$?"
tail -c 1500 gpurun_out/s25_bench_8gpu.json
tail -2 gpurun_out/s25_bench_8gpu.err
