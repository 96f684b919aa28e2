#!/bin/bash
# round 2, thirtieth GPU shot (4 GPUs): smoke() on the final library, then the replica line of bench.py at N = 2 and N = 4
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/s30_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/s30_smoke.log
for n in 2 4; do
  timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29550 + n)) \
    bench.py --gpus $n --steps 20 --warmup 3 --no-e2e --no-c5 --no-cpu-baseline > gpurun_out/s30_bench_${n}gpu.json 2> gpurun_out/s30_bench_${n}gpu.err
  echo "bench --gpus $n rc=$?"
done
