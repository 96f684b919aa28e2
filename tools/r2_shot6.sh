#!/bin/bash
# round 2, sixth GPU shot (2 GPUs): the tests that changed, then bench.py --gpus 2 as the driver launches it (replicas + C5 record)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "c3_full_size or tier3 or reference_run_loop or reference_checkpoint or resume_refuses or set_atoms_keeps" > gpurun_out/s6_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/s6_tests.log
tail -15 gpurun_out/s6_tests.log
nvidia-smi --query-gpu=index,name --format=csv > gpurun_out/s6_gpus.txt
timeout 2400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus 2 --steps 20 --warmup 3 > gpurun_out/s6_bench_2gpu.json 2> gpurun_out/s6_bench_2gpu.err
echo "bench --gpus 2 rc=$?"
tail -c 2500 gpurun_out/s6_bench_2gpu.json
tail -5 gpurun_out/s6_bench_2gpu.err
