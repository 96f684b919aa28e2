"""Regenerates the checkpoint fixtures under tests/golden/.  Run HERE (the build container), where /root/reference
exists:

    make -C oracle ref && python tests/golden/make_checkpoint_fixtures.py

For each case the UNMODIFIED reference (oracle/_ref, scalar build) runs modsimdata with checkpoints enabled
(GapsRunner.cpp:224-256); the last file it leaves behind is committed as ref_checkpoint_<case>.bin, and the outputs
of (a) that uninterrupted run and (b) the reference resuming from the file (GapsRunner.cpp:99-105,258-270) go into
ref_checkpoint_golden.npz.  tests/test_checkpoint.py holds the oracle and the library's reader/writer to these bytes
on machines that have neither /root/reference nor oracle/_ref.
"""
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)

from tests.cases import load_data  # noqa: E402

# name -> (data set, checkpoint interval, run parameters)
CHECKPOINT_CASES = {
    "dense": ("modsim", 25, dict(seed=42, nPatterns=3, nIterations=60, outputFrequency=10)),
    "sparse": ("modsim", 20, dict(seed=7, nPatterns=4, nIterations=50, outputFrequency=10, useSparseOptimization=1)),
}
FIELDS = ("Amean", "Asd", "Pmean", "Psd", "chisqHistory", "atomHistoryA", "atomHistoryP")


def pack(out, prefix, res):
    for f in FIELDS:
        out[prefix + f] = getattr(res, f)
    out[prefix + "scalars"] = np.array([res.totalUpdates, res.meanChiSq, res.averageQueueLengthA,
                                        res.averageQueueLengthP], np.float64)


def main():
    from oracle.harness import RefLib
    ref = RefLib("scalar")
    out = {}
    with tempfile.TemporaryDirectory() as tmp:
        for name, (dataset, interval, kw) in CHECKPOINT_CASES.items():
            data = load_data(dataset)
            path = os.path.join(HERE, "ref_checkpoint_%s.bin" % name)
            full = ref.run(data, checkpointInterval=interval, checkpointOutFile=path, **kw)
            resumed = ref.run(data, checkpointInFile=path, checkpointOutFile=os.path.join(tmp, "again.out"), **kw)
            pack(out, name + "/full/", full)
            pack(out, name + "/resumed/", resumed)
            print(name, os.path.getsize(path), "bytes; meanChiSq", full.meanChiSq, resumed.meanChiSq)
    np.savez_compressed(os.path.join(HERE, "ref_checkpoint_golden.npz"), **out)


if __name__ == "__main__":
    main()
