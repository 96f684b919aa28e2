"""Debug: resident-kernel statistics at the bench workload (COGAPS_PERSISTENT_DEBUG=1 prints per update())."""
import os, sys, time
os.environ["COGAPS_PERSISTENT_DEBUG"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rows, cols = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (20000, 5000)
data = bench.make_data(rows, cols, 20)
chain = bench.Chain(data, 20, 42)
os.environ["COGAPS_PERSISTENT_DEBUG"] = "0"
chain.ramp(int(sys.argv[3]) if len(sys.argv) > 3 else 50)
chain.A.resetCounters(); chain.P.resetCounters()
os.environ["COGAPS_PERSISTENT_DEBUG"] = "1"
t0 = time.perf_counter()
n = 0
for _ in range(5):
    n += chain.step()
dt = time.perf_counter() - t0
os.environ["COGAPS_PERSISTENT_DEBUG"] = "0"
for nm, smp in (("A", chain.A), ("P", chain.P)):
    c = smp.counters()
    print(nm, "batches", c.nBatches, "props", c.nProposalsQueued, "gen s", c.secondsHostGenerate, "wait s", c.secondsDeviceWait,
          "kernel s", c.secondsKernel, "bytes", c.algorithmicBytes)
print("updates/s %.0f  ms/step %.3f" % (n / dt, dt / 5 * 1e3))
