"""Regenerates tests/golden/*.  Run HERE (the build container), where /root/reference exists:

    make -C oracle ref && python tests/golden/make_fixtures.py

It (1) extracts the two data sets the reference ships for this path (data/modsimdata.rda 25x20,
inst/extdata/GIST.csv 1363x9) into .npy, and (2) runs the UNMODIFIED reference (oracle/_ref, scalar
and AVX builds) on a fixed list of cases and stores its outputs as golden vectors, so the oracle can
be pinned on machines where neither /root/reference nor oracle/_ref exists.
"""
import gzip
import os
import struct
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
REF = os.environ.get("COGAPS_REFERENCE", "/root/reference")

from tests.cases import RUN_CASES, load_data  # noqa: E402


def read_modsim():
    """R serialisation (XDR) of a data.frame with 20 REALSXP columns of length 25."""
    b = gzip.decompress(open(os.path.join(REF, "data", "modsimdata.rda"), "rb").read())
    cols, i = [], 0
    while i < len(b) - 8:
        flags, = struct.unpack(">I", b[i:i + 4])
        n, = struct.unpack(">I", b[i + 4:i + 8])
        if (flags & 0xff) == 14 and n == 25 and i + 8 + 8 * n <= len(b):
            cols.append(np.frombuffer(b[i + 8:i + 8 + 8 * n], dtype=">f8").astype(np.float64))
            i += 8 + 8 * n
        else:
            i += 1
    assert len(cols) == 20
    return np.stack(cols, axis=1).astype(np.float32)


def read_gist():
    return np.loadtxt(os.path.join(REF, "inst", "extdata", "GIST.csv"), delimiter=",", skiprows=1,
                      usecols=range(1, 10), dtype=np.float64).astype(np.float32)


def main():
    from oracle.harness import RefLib
    modsim, gist = read_modsim(), read_gist()
    assert modsim.shape == (25, 20) and gist.shape == (1363, 9)
    np.save(os.path.join(HERE, "modsim.npy"), modsim)
    np.save(os.path.join(HERE, "gist.npy"), gist)
    datasets = {"modsim": modsim, "gist": gist}
    out = {}
    for variant in ("scalar", "avx"):
        ref = RefLib(variant)
        erf, erfinv, qgamma = ref.tables()
        out["%s/tables/erf" % variant] = erf
        out["%s/tables/erfinv" % variant] = erfinv
        out["%s/tables/qgamma" % variant] = qgamma
        for name, case in RUN_CASES.items():
            data = datasets[case["data"]] if case["data"] in datasets else load_data(case["data"])
            kw = dict(case["params"])
            unc = None
            if case.get("uncertainty"):
                unc = np.maximum(0.15 * data, 0.2).astype(np.float32)
            if case.get("fixed"):
                rows = data.shape[1] if kw["whichMatrixFixed"] == "P" else data.shape[0]
                if kw.get("transposeData"):
                    rows = data.shape[0] if kw["whichMatrixFixed"] == "P" else data.shape[1]
                rng = np.random.default_rng(7)
                kw["fixedPatterns"] = rng.gamma(2.0, 0.5, (rows, kw["nPatterns"])).astype(np.float32)
            res = ref.run(data, uncertainty=unc, snapshots=True, **kw)
            pre = "%s/%s/" % (variant, name)
            for f in ("Amean", "Asd", "Pmean", "Psd", "chisqHistory", "atomHistoryA", "atomHistoryP"):
                out[pre + f] = getattr(res, f)
            out[pre + "scalars"] = np.array([res.totalUpdates, res.meanChiSq, res.averageQueueLengthA,
                                             res.averageQueueLengthP], np.float64)
            if case.get("pump"):
                out[pre + "pumpMatrix"] = res.pumpMatrix
                out[pre + "meanPatternAssignment"] = res.meanPatternAssignment
            if kw.get("snapshotFrequency"):
                out[pre + "snapshotsA_last"] = res.snapshotsA[-1]
                out[pre + "snapshotsP_last"] = res.snapshotsP[-1]
            print(variant, name, res.totalUpdates, res.atomHistoryA[-1], res.atomHistoryP[-1], res.meanChiSq)
    np.savez_compressed(os.path.join(HERE, "ref_golden.npz"), **out)


if __name__ == "__main__":
    main()
