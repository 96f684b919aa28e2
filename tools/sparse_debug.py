"""Debug: resident-kernel statistics of the sparse sampler (COGAPS_PERSISTENT_DEBUG=1 prints per update())."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
rows, cols, k = (int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (20000, 12000, 50)
data = bench.make_data(rows, cols, k, zero_fraction=0.95)
chain = bench.Chain(data, k, 42, sparse=True)
chain.ramp(30)
chain.A.resetCounters(); chain.P.resetCounters()
os.environ["COGAPS_PERSISTENT_DEBUG"] = "1"
t0 = time.perf_counter()
n = 0
for _ in range(3):
    n += chain.step()
dt = time.perf_counter() - t0
os.environ["COGAPS_PERSISTENT_DEBUG"] = "0"
for nm, smp in (("A", chain.A), ("P", chain.P)):
    c = smp.counters()
    print(nm, "atoms", smp.nAtoms(), "batches", c.nBatches, "props", c.nProposalsQueued, "gen s", c.secondsHostGenerate, "wait s", c.secondsDeviceWait,
          "kernel s", c.secondsKernel)
print("updates/s %.0f  ms/step %.3f" % (n / dt, dt / 3 * 1e3))
