"""Debug: average per-phase timeline of the eval kernel's leader CTAs at the bench workload."""
import ctypes as C, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import bench
import cogaps_b200 as cg
from cogaps_b200._lib import check
L = cg.lib()
L.cgb_sampler_debug_phase_clocks.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_double), C.POINTER(C.c_uint64)]
rows, cols = int(sys.argv[1]) if len(sys.argv) > 1 else 20000, int(sys.argv[2]) if len(sys.argv) > 2 else 5000
data = bench.make_data(rows, cols, 20)
chain = bench.Chain(data, 20, 42)
chain.ramp(int(sys.argv[3]) if len(sys.argv) > 3 else 40)
names = ["start-spread(ns)/launches", "", "copies issued", "data landed", "scan done", "cluster reduce", "decision", "broadcast", "commit", "exit"]
for smp, nm in ((chain.A, "A"), (chain.P, "P")):
    check(L.cgb_sampler_debug_phase_clocks(smp._h, 1, None, None))
chain.step(); chain.step()
for smp, nm in ((chain.A, "A"), (chain.P, "P")):
    out = (C.c_double * 12)(); n = C.c_uint64()
    check(L.cgb_sampler_debug_phase_clocks(smp._h, 0, out, C.byref(n)))
    print(nm, "tasks", n.value, "launches", out[1], "start spread ns/launch", out[0] / max(out[1], 1))
    for i in range(2, 9):
        print("   %-16s %8.0f cycles" % (names[i], out[i] / max(n.value, 1)))
