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
    ap.add_argument("--configs", default="", help="';'-separated sets of NAME=VALUE,... environment knobs, each timed in turn on the same chain "
                                                  "(the library reads its COGAPS_SWEEP_* knobs at every update)")
    args = ap.parse_args()
    import torch
    data = bench.make_data(args.genes, args.samples, args.patterns)
    chain = bench.Chain(data, args.patterns, 42, updateMode=args.mode)
    t0 = time.time()
    chain.ramp(args.ramp)
    torch.cuda.synchronize()
    tr = time.time() - t0
    import numpy as np
    for name, smp, nRows in (("A", chain.A, args.genes), ("P", chain.P, args.samples)):
        pos, _ = smp.atoms()
        binLength = 0xFFFFFFFFFFFFFFFF // (nRows * args.patterns)
        rows = np.minimum(np.asarray(pos, dtype=np.uint64) // np.uint64(binLength * args.patterns), np.uint64(nRows - 1)).astype(np.int64)
        cnt = np.bincount(rows, minlength=nRows)
        print(json.dumps({"atoms_per_row": name, "mean": float(cnt.mean()), "p50": int(np.percentile(cnt, 50)), "p90": int(np.percentile(cnt, 90)),
                          "p99": int(np.percentile(cnt, 99)), "p999": int(np.percentile(cnt, 99.9)), "max": int(cnt.max()), "empty_rows": int((cnt == 0).sum())}), flush=True)
    configs = [c for c in args.configs.split(";")] if args.configs else [""]
    for cfg in configs:
        knobs = dict(kv.split("=") for kv in cfg.split(",") if kv)
        for k, v in knobs.items():
            os.environ[k] = v
        chain.step()  # first use of a kernel instance (module load, shared-memory opt-in) stays outside the timing
        torch.cuda.synchronize()
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
            "config": cfg, "mode": "sweep" if args.mode else "exact", "shape": [args.genes, args.samples, args.patterns],
            "ramp_iters": args.ramp, "ramp_s": round(tr, 3), "steps": args.steps, "ms_per_step": round(1e3 * dt / args.steps, 3),
            "updates_per_s": round(made / dt), "asked": n, "made": made,
            "atomsA": chain.A.nAtoms(), "atomsP": chain.P.nAtoms(), "chisq": chain.P.chiSq(),
            "kernel_ms_per_step_A": round(1e3 * cA.secondsKernel / args.steps, 3),
            "kernel_ms_per_step_P": round(1e3 * cP.secondsKernel / args.steps, 3),
            "algorithmic_GBps_A": round(cA.algorithmicBytes / max(cA.secondsKernel, 1e-9) / 1e9, 1),
            "algorithmic_GBps_P": round(cP.algorithmicBytes / max(cP.secondsKernel, 1e-9) / 1e9, 1),
            "scans_A": cA.nProposalsQueued, "scans_P": cP.nProposalsQueued,
        }
        print(json.dumps(out), flush=True)
        for k in knobs:
            del os.environ[k]


if __name__ == "__main__":
    main()
