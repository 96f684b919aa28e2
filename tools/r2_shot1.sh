#!/bin/bash
# round 2, first GPU shot: sweep parity tests, sweep timing at C3, launch list + one full ncu capture of the sweep kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm --format=csv > gpurun_out/s1_gpu.txt 2>&1
timeout 900 python -m pytest tests/test_sweep.py -m gpu -x -q > gpurun_out/s1_sweep_tests.log 2>&1
echo "sweep tests rc=$?" >> gpurun_out/s1_sweep_tests.log
tail -5 gpurun_out/s1_sweep_tests.log
timeout 600 python tools/sweep_bench.py --ramp 200 --steps 20 > gpurun_out/s1_sweep_bench.json 2> gpurun_out/s1_sweep_bench.err
cat gpurun_out/s1_sweep_bench.json
timeout 600 python tools/sweep_bench.py --ramp 600 --steps 20 >> gpurun_out/s1_sweep_bench.json 2>> gpurun_out/s1_sweep_bench.err
tail -1 gpurun_out/s1_sweep_bench.json
timeout 300 python tools/sweep_bench.py --mode 0 --ramp 50 --steps 10 >> gpurun_out/s1_sweep_bench.json 2>> gpurun_out/s1_sweep_bench.err
tail -1 gpurun_out/s1_sweep_bench.json
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/s1_sweep_launches.csv \
  python tools/sweep_bench.py --ramp 60 --steps 3 > gpurun_out/s1_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_kernel -s 100 -c 2 -o gpurun_out/s1_sweep_full -f \
  python tools/sweep_bench.py --ramp 60 --steps 3 > gpurun_out/s1_ncu_full.log 2>&1
ls -la gpurun_out | tail -20
