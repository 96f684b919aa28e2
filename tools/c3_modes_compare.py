"""Exact mode and sweep mode side by side at BASELINE configs[2]'s size (20000 x 5000, k=20): atom-count and chi-square
trajectories of a short run from the same seed.  Prints one JSON line per mode."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402
import cogaps_b200 as cg  # noqa: E402

its = int(sys.argv[1]) if len(sys.argv) > 1 else 200
data = bench.make_data()
for mode in (1, 0):
    t0 = time.perf_counter()
    r = cg.gaps_run(data, seed=42, nPatterns=bench.K, nIterations=its, outputFrequency=max(its // 4, 1), maxThreads=1, updateMode=mode)
    print(json.dumps({"mode": "sweep" if mode else "exact", "iterations_per_phase": its, "wall_s": time.perf_counter() - t0,
                      "atoms_A": [int(x) for x in r.atomHistoryA], "atoms_P": [int(x) for x in r.atomHistoryP],
                      "chisq": [float(x) for x in r.chisqHistory], "meanChiSq": float(r.meanChiSq), "updates": int(r.totalUpdates)}), flush=True)
