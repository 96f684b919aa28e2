#!/bin/bash
# round 2, eighteenth GPU shot (2 GPUs): bench.py --gpus 2 as the driver launches it, after the C5 gather moved to the sparse model's row copy
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "comm_allgather" > gpurun_out/s18_comm_test.log 2>&1; tail -2 gpurun_out/s18_comm_test.log
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s18_bench_2gpu.json 2> gpurun_out/s18_bench_2gpu.err
echo "bench --gpus 2 rc=$?"
tail -c 1500 gpurun_out/s18_bench_2gpu.json
tail -3 gpurun_out/s18_bench_2gpu.err
