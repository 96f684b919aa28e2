#!/bin/bash
# round 2, seventeenth GPU shot (8 GPUs): dist_check at 2 / 4 / 8 ranks (distributed result == single-process result, bit for bit),
# then bench.py --gpus 8 as the driver launches it (replica line + the C5 record with the all-gather over 8 ranks)
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/s17_gpus.txt
for n in 2 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29540 + n)) \
    tools/dist_check.py > gpurun_out/s17_dist_check_$n.log 2>&1
  echo "dist_check $n rc=$? $(grep dist_check gpurun_out/s17_dist_check_$n.log | tail -1)"
done
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 8 --steps 20 --warmup 3 > gpurun_out/s17_bench_8gpu.json 2> gpurun_out/s17_bench_8gpu.err
echo "bench --gpus 8 rc=$?"
tail -c 1800 gpurun_out/s17_bench_8gpu.json
tail -3 gpurun_out/s17_bench_8gpu.err
