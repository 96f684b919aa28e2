#!/bin/bash
# round 2, twentieth GPU shot (8 GPUs): the replica line of bench.py --gpus 8 with one clock sampler for the whole job (rank 0) instead of one per rank
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e --no-c5 --no-cpu-baseline > gpurun_out/s20_bench_8gpu.json 2> gpurun_out/s20_bench_8gpu.err
echo "rc=$?"; tail -c 900 gpurun_out/s20_bench_8gpu.json; tail -2 gpurun_out/s20_bench_8gpu.err
