#!/bin/bash
# round 2, eighth GPU shot: sparse sweep parity, sparse exact regression (scan refactor), C4 in both modes
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_sweep.py -m gpu -x -q > gpurun_out/s8_sweep_tests.log 2>&1
tail -12 gpurun_out/s8_sweep_tests.log
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "sparse" > gpurun_out/s8_sparse_exact_tests.log 2>&1
tail -3 gpurun_out/s8_sparse_exact_tests.log
timeout 1500 python - > gpurun_out/s8_c4.json 2> gpurun_out/s8_c4.err <<'PY'
import json, sys, time
sys.path.insert(0, '.')
import bench, torch
g, s, k = 50000, 30000, 50
data = bench.make_data(g, s, k, bench.DATA_SEED, 0.95)
for mode, ramp, steps in ((1, 150, 10), (0, 30, 5)):
    chain = bench.Chain(data, k, 42, sparse=True, updateMode=mode)
    t0 = time.time(); chain.ramp(ramp); torch.cuda.synchronize(); tr = time.time() - t0
    for smp in (chain.A, chain.P): smp.resetCounters()
    t0 = time.time(); n = 0
    for _ in range(steps): n += chain.step()
    torch.cuda.synchronize(); dt = time.time() - t0
    cA, cP = chain.A.counters(), chain.P.counters()
    made = cA.nProposalsTotal + cP.nProposalsTotal
    print(json.dumps({"workload": "sparse 50000x30000 95% zeros k=50", "mode": "sweep" if mode else "exact", "ramp_iters": ramp, "ramp_s": tr,
                      "steps": steps, "ms_per_step": 1e3 * dt / steps, "updates_per_s": made / dt, "atomsA": chain.A.nAtoms(), "atomsP": chain.P.nAtoms(),
                      "kernel_ms_per_step_A": 1e3 * cA.secondsKernel / steps, "kernel_ms_per_step_P": 1e3 * cP.secondsKernel / steps}), flush=True)
    del chain
PY
cat gpurun_out/s8_c4.json; tail -3 gpurun_out/s8_c4.err
