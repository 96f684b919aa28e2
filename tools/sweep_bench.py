"""Sweep-mode timing at a BASELINE shape (default C3: 20000x5000 k=20): ramp, then time whole iterations.
Prints one JSON line per measurement.  Scratch tool for development; bench.py carries the judged numbers."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--genes", type=int, default=20000)
    ap.add_argument("--samples", type=int, default=5000)
    ap.add_argument("--patterns", type=int, default=20)
    ap.add_argument("--ramp", type=int, default=200)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--mode", type=int, default=1)
    args = ap.parse_args()
    import torch
    data = bench.make_data(args.genes, args.samples, args.patterns)
    chain = bench.Chain(data, args.patterns, 42, updateMode=args.mode)
    t0 = time.time()
    chain.ramp(args.ramp)
    torch.cuda.synchronize()
    tr = time.time() - t0
    for smp in (chain.A, chain.P):
        smp.resetCounters()
    t0 = time.time()
    n = 0
    for _ in range(args.steps):
        n += chain.step()
    torch.cuda.synchronize()
    dt = time.time() - t0
    cA, cP = chain.A.counters(), chain.P.counters()
    made = cA.nProposalsTotal + cP.nProposalsTotal
    out = {
        "mode": "sweep" if args.mode else "exact", "shape": [args.genes, args.samples, args.patterns],
        "ramp_iters": args.ramp, "ramp_s": round(tr, 3), "steps": args.steps, "ms_per_step": round(1e3 * dt / args.steps, 3),
        "updates_per_s": round(made / dt), "asked": n, "made": made,
        "atomsA": chain.A.nAtoms(), "atomsP": chain.P.nAtoms(), "chisq": chain.P.chiSq(),
        "kernel_ms_per_step_A": round(1e3 * cA.secondsKernel / args.steps, 3),
        "kernel_ms_per_step_P": round(1e3 * cP.secondsKernel / args.steps, 3),
        "algorithmic_GBps_A": round(cA.algorithmicBytes / max(cA.secondsKernel, 1e-9) / 1e9, 1),
        "algorithmic_GBps_P": round(cP.algorithmicBytes / max(cP.secondsKernel, 1e-9) / 1e9, 1),
        "scans_A": cA.nProposalsQueued, "scans_P": cP.nProposalsQueued,
    }
    print(json.dumps(out))


if __name__ == "__main__":
    main()
