"""Debug: where the end-to-end cgb_run call of bench.py (100 iterations per phase from zero atoms, sweep mode) spends its wall time."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import cogaps_b200 as cg
data = bench.make_data()
os.environ['COGAPS_HOST_PROFILE'] = '1'
for it in (100, 100, 100, 100, 100):
    t0 = time.perf_counter()
    res = cg.gaps_run(data, seed=42, nPatterns=20, nIterations=it, outputFrequency=0, maxThreads=1, updateMode=1)
    wall = time.perf_counter() - t0
    print("wall %.3f s | run loop %.3f s (A.update %.3f, P.update %.3f, device kernels %.3f) | outside the loop %.3f s | updates %d"
          % (wall, res.totalRunningTime, res.secondsUpdateA, res.secondsUpdateP, res.secondsDevice, wall - res.totalRunningTime, res.totalUpdates), flush=True)
