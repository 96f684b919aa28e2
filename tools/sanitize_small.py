"""A small sweep + exact run for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): modsimdata, a few iterations
of every sampler; the exact path uses one launch per batch under the tool."""
import os
import sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import cogaps_b200 as cg  # noqa: E402

data = np.load(os.path.join(ROOT, "tests", "golden", "modsim.npy"))
gist = np.load(os.path.join(ROOT, "tests", "golden", "gist.npy"))[:200]
its = int(sys.argv[1]) if len(sys.argv) > 1 else 12
for d, k in ((data, 3), (gist, 4)):
    for extra in (dict(updateMode=1), dict(updateMode=1, useSparseOptimization=1), dict(), dict(useSparseOptimization=1)):
        r = cg.gaps_run(d, seed=3, nPatterns=k, nIterations=its, outputFrequency=its, maxThreads=1, **extra)
        print(d.shape, extra, "updates", r.totalUpdates, "atoms", r.atomHistoryA[-1], r.atomHistoryP[-1], flush=True)
# rows beyond 10240 floats: 512 threads, only the AP line in shared memory (STAGE 2); 200 rows: handed out longest chain first
from tests.cases import load_data  # noqa: E402
long_rows = load_data("syn:20:12000:3:5")
r = cg.gaps_run(long_rows, seed=3, nPatterns=3, nIterations=max(its // 3, 2), outputFrequency=its, maxThreads=1, updateMode=1)
print(long_rows.shape, "sweep, long rows", r.totalUpdates, flush=True)
unc = np.maximum(0.15 * data, 0.2).astype(np.float32)
r = cg.gaps_run(data, uncertainty=unc, seed=3, nPatterns=3, nIterations=its, outputFrequency=its, maxThreads=1, updateMode=1)
print("explicit uncertainty", r.totalUpdates, flush=True)
