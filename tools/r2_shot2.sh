#!/bin/bash
# round 2, second GPU shot: sweep v2 (counter-based draws, transport), whole GPU suite, sweep timing, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep.py -m gpu -x -q > gpurun_out/s2_sweep_tests.log 2>&1
echo "sweep tests rc=$?" >> gpurun_out/s2_sweep_tests.log
tail -4 gpurun_out/s2_sweep_tests.log
timeout 600 python tools/sweep_bench.py --ramp 600 --steps 20 > gpurun_out/s2_sweep_bench.json 2> gpurun_out/s2_sweep_bench.err
tail -1 gpurun_out/s2_sweep_bench.json
COGAPS_SWEEP_TRANSPORT=0 timeout 600 python tools/sweep_bench.py --ramp 600 --steps 20 >> gpurun_out/s2_sweep_bench.json 2>> gpurun_out/s2_sweep_bench.err
tail -1 gpurun_out/s2_sweep_bench.json
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sweep_ -s 1000 -c 4 -o gpurun_out/s2_sweep_full -f \
  python tools/sweep_bench.py --ramp 260 --steps 3 > gpurun_out/s2_ncu_full.log 2>&1
timeout 1500 python -m pytest tests -m gpu -x -q --deselect tests/test_sweep.py > gpurun_out/s2_gpu_tests.log 2>&1
echo "gpu tests rc=$?" >> gpurun_out/s2_gpu_tests.log
tail -4 gpurun_out/s2_gpu_tests.log
