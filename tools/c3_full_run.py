"""BASELINE.json configs[2] in full: synthetic dense 20000x5000, nPatterns=20, 10k iterations per phase, through cgb_run
(= gaps::run) on host buffers.  Prints one JSON line: wall time, atom-updates/s, the chi-square and atom-count
trajectories (every 1000 iterations) and meanChiSq.  Default: the row-parallel sweep; --mode 0 runs the reference's own
chain (hours at this size: use --iterations to bound it)."""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cogaps_b200 as cg  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iterations", type=int, default=10000)
    ap.add_argument("--mode", type=int, default=1)
    ap.add_argument("--seed", type=int, default=42)
    args = ap.parse_args()
    data = bench.make_data()
    t0 = time.perf_counter()
    res = cg.gaps_run(data, seed=args.seed, nPatterns=bench.K, nIterations=args.iterations,
                      outputFrequency=max(args.iterations // 10, 1), maxThreads=1, updateMode=args.mode)
    wall = time.perf_counter() - t0
    print(json.dumps({
        "workload": "synthetic dense %dx%d nPatterns=%d, %d iterations per phase (BASELINE.json configs[2])" % (bench.G, bench.S, bench.K, args.iterations),
        "update_mode": "sweep" if args.mode else "exact", "seed": args.seed, "wall_s": wall,
        "sampler_loop_s": float(res.totalRunningTime), "atom_updates": int(res.totalUpdates),
        "atom_updates_per_s_whole_call": res.totalUpdates / wall,
        "atom_updates_per_s_sampler_loop": res.totalUpdates / max(float(res.totalRunningTime), 1e-9),
        "ms_per_iteration": float(res.totalRunningTime) / (2 * args.iterations) * 1e3,
        "chisq_history": [float(x) for x in res.chisqHistory], "atoms_A_history": [int(x) for x in res.atomHistoryA],
        "atoms_P_history": [int(x) for x in res.atomHistoryP], "meanChiSq": float(res.meanChiSq),
        "chisq_per_element_final": float(res.chisqHistory[-1]) / (bench.G * bench.S),
        "seconds_update_A": float(res.secondsUpdateA), "seconds_update_P": float(res.secondsUpdateP),
        "seconds_device_kernels": float(res.secondsDevice), "algorithmic_bytes": float(res.algorithmicBytes)}))


if __name__ == "__main__":
    main()
