"""Shared plumbing for calling a `*_run(data, nrow, ncol, [colmajor,] unc, params*, result*)` entry
point through ctypes: allocates the caller-owned result arrays and wraps them as numpy arrays."""
import ctypes as C
import numpy as np
from ._abi import CgbParams, CgbResult, c_float_p, c_u32_p


def fptr(a):
    return a.ctypes.data_as(c_float_p) if a is not None else None


def make_params(**kw):
    p = CgbParams.defaults()
    keep = []
    for k, v in kw.items():
        if k == "whichMatrixFixed" and isinstance(v, str):
            v = ord(v)
        if k == "subsetIndices" and v is not None:
            arr = np.ascontiguousarray(v, dtype=np.uint32)
            keep.append(arr)
            p.subsetIndices = arr.ctypes.data_as(c_u32_p)
            p.nSubsetIndices = arr.size
            continue
        if k == "fixedPatterns" and v is not None:
            arr = np.ascontiguousarray(v, dtype=np.float32)
            keep.append(arr)
            p.fixedPatterns = arr.ctypes.data_as(c_float_p)
            continue
        if not hasattr(p, k):
            raise TypeError("unknown parameter %r" % k)
        setattr(p, k, v)
    p._keepalive = keep
    return p


def result_dims(params, nrow, ncol):
    g, s = (ncol, nrow) if params.transposeData else (nrow, ncol)
    if params.nSubsetIndices:
        if params.subsetGenes:
            g = params.nSubsetIndices
        else:
            s = params.nSubsetIndices
    return g, s


class ResultArrays(object):
    """Owns the numpy arrays a cgb_result points into."""

    def __init__(self, params, nrow, ncol, snapshots=False):
        g, s = result_dims(params, nrow, ncol)
        k = params.nPatterns
        self.nGenes, self.nSamples, self.nPatterns = g, s, k
        nhist = 0
        if params.outputFrequency:
            nhist = 2 * (params.nIterations // params.outputFrequency) + 2
        self.Amean = np.zeros((g, k), np.float32)
        self.Asd = np.zeros((g, k), np.float32)
        self.Pmean = np.zeros((s, k), np.float32)
        self.Psd = np.zeros((s, k), np.float32)
        self.chisqHistory = np.zeros(max(nhist, 1), np.float32)
        self.atomHistoryA = np.zeros(max(nhist, 1), np.uint32)
        self.atomHistoryP = np.zeros(max(nhist, 1), np.uint32)
        self.pumpMatrix = np.zeros((g, k), np.float32)
        self.meanPatternAssignment = np.zeros((g, k), np.float32)
        nsnap = 0
        if snapshots and params.snapshotFrequency:
            nsnap = 2 * (params.nIterations // params.snapshotFrequency)
        self.snapshotsA = np.zeros((max(nsnap, 1), g, k), np.float32)
        self.snapshotsP = np.zeros((max(nsnap, 1), s, k), np.float32)
        r = CgbResult()
        r.struct_size = C.sizeof(CgbResult)
        r.historyCapacity = nhist
        r.Amean, r.Asd, r.Pmean, r.Psd = map(fptr, (self.Amean, self.Asd, self.Pmean, self.Psd))
        r.chisqHistory = fptr(self.chisqHistory)
        r.atomHistoryA = self.atomHistoryA.ctypes.data_as(c_u32_p)
        r.atomHistoryP = self.atomHistoryP.ctypes.data_as(c_u32_p)
        r.pumpMatrix = fptr(self.pumpMatrix)
        r.meanPatternAssignment = fptr(self.meanPatternAssignment)
        r.snapshotsA = fptr(self.snapshotsA) if nsnap else None
        r.snapshotsP = fptr(self.snapshotsP) if nsnap else None
        r.snapshotCapacity = nsnap
        self.c = r

    def finish(self):
        r = self.c
        n = r.nHistory
        self.chisqHistory = self.chisqHistory[:n].copy()
        self.atomHistoryA = self.atomHistoryA[:n].copy()
        self.atomHistoryP = self.atomHistoryP[:n].copy()
        ns = r.nSnapshotsEquilibration + r.nSnapshotsSampling
        self.snapshotsA = self.snapshotsA[:ns]
        self.snapshotsP = self.snapshotsP[:ns]
        for f in ("nSnapshotsEquilibration", "nSnapshotsSampling", "seed", "totalUpdates",
                  "totalRunningTime", "meanChiSq", "averageQueueLengthA", "averageQueueLengthP",
                  "nBatchesA", "nBatchesP", "secondsUpdateA", "secondsUpdateP", "secondsDevice",
                  "algorithmicBytes"):
            setattr(self, f, getattr(r, f))
        return self
