"""Debug: three whole runs from three threads sharing one GPU (the opt-in stress test), with progress output."""
import faulthandler, os, sys, threading, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
faulthandler.dump_traceback_later(45, exit=True)
import numpy as np
import cogaps_b200 as cg
from cogaps_b200._lib import check
from tests.test_gpu_parity import case_inputs
cases = sys.argv[1:] or ["gist_async", "syn_203x117", "sparse_120x90"]
inputs = [case_inputs(n, nIterations=40) for n in cases]
check(cg.lib().cgb_set_resident_share(len(cases)))
done = [None] * len(cases)
def drive(i):
    d, _, kw = inputs[i]
    kw = dict(kw, printMessages=1)
    t0 = time.time()
    done[i] = cg.gaps_run(d, **kw)
    print("chain %d (%s) finished in %.2f s" % (i, cases[i], time.time() - t0), flush=True)
ts = [threading.Thread(target=drive, args=(i,)) for i in range(len(cases))]
for t in ts: t.start()
for t in ts: t.join()
print("all done", [r is not None for r in done])
