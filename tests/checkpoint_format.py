"""An independent reader of the reference's checkpoint wire format, for the tests.

The format is what `Archive` (src/utils/Archive.h:16-87: raw little-endian scalars behind a u32 magic number) receives
from createCheckpoint (src/GapsRunner.cpp:237-240):

    magic  params  randState  ASampler  PSampler  stats  int(phase)  iter  rng

with, field by field,
    params     GapsParameters.cpp:82-88    seed nGenes nSamples nPatterns nIterations (u32) alphaA alphaP maxGibbsMassA
                                           maxGibbsMassP (f32) useSparseOptimization (bool, 1 byte) checkpointInterval (u32)
    randState  Random.cpp:250-254,347-351  the xoroshiro128+ seeder: two u64
    sampler    AsynchronousGibbsSampler.h:221-226   model, domain, queue
      dense model   DenseNormalModel.cpp:260-264 -> Matrix.cpp:182-190: nRows nCols (u32), per column a Vector
                    (Vector.cpp:90-98: size (u32) then the floats)
      sparse model  SparseNormalModel.cpp:313-317 -> HybridMatrix.cpp:85-97: nRows nCols, every row as a Vector, every
                    column as a HybridVector (HybridVector.cpp:103-115: size, size/64+1 u64 flag words, the floats); beta
      domain        ConcurrentAtomicDomain.cpp:134-142: domainLength (u64), atom count (u64), atoms in pick-vector order
                    as (pos u64, mass f32) (ConcurrentAtom.cpp:98-102)
      queue         ProposalQueue.cpp:285-291: rng state, minAtoms, maxAtoms, binLength, numCols (u64) alpha domainLength
                    numBins (f64) lambda (f32) useCachedRng (bool) u1 u2 (f32)
    stats      GapsStatistics.cpp:164-169: the four running-sum matrices, statUpdates, numPatterns (u32)
    phase (i32), iter (u32), rng (u64)     GapsRunner.cpp:240, Random.cpp:202-206

This file is test infrastructure; the product's reader/writer is cogaps_b200/csrc/checkpoint.cpp.
"""
import struct

import numpy as np

MAGIC = 0xB123AA4D  # utils/Archive.h:16


class _Cursor(object):
    def __init__(self, raw):
        self.raw = raw
        self.off = 0

    def take(self, fmt):
        size = struct.calcsize("<" + fmt)
        vals = struct.unpack_from("<" + fmt, self.raw, self.off)
        self.off += size
        return vals if len(vals) > 1 else vals[0]

    def array(self, dtype, n):
        dt = np.dtype(dtype).newbyteorder("<")
        a = np.frombuffer(self.raw, dtype=dt, count=n, offset=self.off)
        self.off += dt.itemsize * n
        return a.copy()


def _vector(c):
    n = c.take("I")
    return c.array(np.float32, n)


def _matrix(c):
    nr, nc = c.take("II")
    cols = [_vector(c) for _ in range(nc)]
    assert all(v.size == nr for v in cols)
    return np.stack(cols, axis=1) if nc else np.zeros((nr, 0), np.float32)   # rows x patterns


def _hybrid_matrix(c):
    nr, nc = c.take("II")
    rows = np.stack([_vector(c) for _ in range(nr)], axis=0)
    cols, flags = [], []
    for _ in range(nc):
        n = c.take("I")
        flags.append(c.array(np.uint64, n // 64 + 1))
        cols.append(c.array(np.float32, n))
    return rows, np.stack(cols, axis=1), flags


def _sampler(c, sparse):
    s = {}
    if sparse:
        s["rows"], s["cols"], s["flags"] = _hybrid_matrix(c)
        s["matrix"] = s["rows"]
        s["beta"] = c.take("f")
    else:
        s["matrix"] = _matrix(c)
    s["domainLength"], n = c.take("QQ")
    atoms = c.array(np.dtype([("pos", "<u8"), ("mass", "<f4")]), n)
    s["pos"], s["mass"] = atoms["pos"].copy(), atoms["mass"].copy()
    (s["rng"], s["minAtoms"], s["maxAtoms"], s["binLength"], s["numCols"], s["alpha"], s["queueDomainLength"],
     s["numBins"], s["lambda"], s["useCachedRng"], s["u1"], s["u2"]) = c.take("QQQQQdddf?ff")
    return s


def parse(raw):
    """bytes of a checkpoint file -> dict; raises if the file is not consumed exactly."""
    c = _Cursor(raw)
    out = {"magic": c.take("I")}
    assert out["magic"] == MAGIC, hex(out["magic"])
    p = {}
    (p["seed"], p["nGenes"], p["nSamples"], p["nPatterns"], p["nIterations"], p["alphaA"], p["alphaP"],
     p["maxGibbsMassA"], p["maxGibbsMassP"], p["useSparseOptimization"], p["checkpointInterval"]) = c.take("IIIIIffff?I")
    out["params"] = p
    out["seeder"] = c.take("QQ")
    out["A"] = _sampler(c, p["useSparseOptimization"])
    out["P"] = _sampler(c, p["useSparseOptimization"])
    out["AmeanSum"], out["AsqSum"], out["PmeanSum"], out["PsqSum"] = (_matrix(c) for _ in range(4))
    out["statUpdates"], out["statPatterns"] = c.take("II")
    out["phase"], out["iter"], out["rng"] = c.take("iIQ")
    assert c.off == len(raw), (c.off, len(raw))
    return out


def parse_file(path):
    with open(path, "rb") as f:
        return parse(f.read())
