"""Regenerates tests/golden/tier3_midsize_ref.npz.  Run HERE (the build container), where /root/reference exists:

    make -C oracle ref && python tests/golden/make_tier3_midsize.py

The UNMODIFIED reference (oracle/_ref, scalar build) on a mid-size synthetic matrix (1200 x 500, 8 patterns:
tests/cases.py "syn:1200:500:8:5") over six seeds, 300 + 300 iterations: atom-count and chi-square trajectories every
100 iterations and meanChiSq.  About 20 s per seed on one core — too long to repeat inside every GPU test run, so the
trajectories travel as a fixture and tests/test_gpu_parity.py::test_tier3_midsize_chains_agree_with_the_reference holds
the free-running CUDA chains (exact mode and sweep) to them.
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from oracle.harness import RefLib  # noqa: E402
from tests.cases import load_data  # noqa: E402

SPEC, K, ITS, FREQ = "syn:1200:500:8:5", 8, 300, 100
SEEDS = [1, 3, 5, 7, 9, 11]


def main():
    ref = RefLib("scalar")
    data = load_data(SPEC)
    rows = []
    for seed in SEEDS:
        r = ref.run(data, seed=seed, nPatterns=K, nIterations=ITS, outputFrequency=FREQ)
        rows.append(np.concatenate([r.atomHistoryA, r.atomHistoryP, r.chisqHistory, [r.meanChiSq]]).astype(np.float64))
        print("seed %d: atoms A %s P %s chi-square %s meanChiSq %.1f" % (seed, r.atomHistoryA[-1], r.atomHistoryP[-1], r.chisqHistory[-1], r.meanChiSq))
    np.savez(os.path.join(HERE, "tier3_midsize_ref.npz"), rows=np.array(rows), seeds=np.array(SEEDS), spec=SPEC, k=K, its=ITS, freq=FREQ)


if __name__ == "__main__":
    main()
