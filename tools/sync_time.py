"""Times sync (the AP transpose) and chi-square at BASELINE configs[2]'s size; prints milliseconds and fractions of the measured HBM peak."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import bench
peak, _ = bench.load_peaks()
data = bench.make_data()
chain = bench.Chain(data, bench.K, 42, updateMode=1)
chain.ramp(30)
def wall_ms(fn, n=20):
    fn(); torch.cuda.synchronize()
    t = time.perf_counter()
    for _ in range(n):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t) / n * 1e3
nbytes = 8.0 * bench.G * bench.S
for name, fn in (("A.sync(P)", lambda: chain.A.sync(chain.P)), ("P.sync(A)", lambda: chain.P.sync(chain.A)), ("P.chiSq", chain.P.chiSq)):
    ms = wall_ms(fn)
    print("%s: %.3f ms = %.2f of the HBM peak (%.0f GB/s)" % (name, ms, nbytes / (ms * 1e-3) / 1e9 / peak, peak))
