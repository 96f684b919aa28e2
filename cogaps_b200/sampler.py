"""Object wrappers over the sampler-level C ABI — the reference's Sampler concept
(AsynchronousGibbsSampler<DenseNormalModel>, GapsRandomState, GapsStatistics) with the same method names,
so tests read like the reference's own C++ tests."""
import ctypes as C

import numpy as np

from ._abi import (CgbSamplerCounters, CgbReductionOrder, c_float_p, c_u32_p, c_u64_p, c_i32_p,
                   ERF_TABLE_SIZE, ERFINV_TABLE_SIZE, QGAMMA_TABLE_SIZE)
from ._lib import lib, check
from ._runhelp import make_params, fptr


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


class GapsRandomState(object):
    """math/Random.h:79-98"""

    def __init__(self, seed, tables=None):
        self._h = C.c_void_p()
        check(lib().cgb_randstate_create(C.c_uint32(seed), C.byref(self._h)))
        if tables is not None:
            erf, erfinv, qgamma = (_f32(t) for t in tables)
            check(lib().cgb_randstate_set_tables(self._h, fptr(erf), fptr(erfinv), fptr(qgamma)))

    def tables(self):
        erf = np.zeros(ERF_TABLE_SIZE, np.float32)
        erfinv = np.zeros(ERFINV_TABLE_SIZE, np.float32)
        qgamma = np.zeros(QGAMMA_TABLE_SIZE, np.float32)
        check(lib().cgb_randstate_get_tables(self._h, fptr(erf), fptr(erfinv), fptr(qgamma)))
        return erf, erfinv, qgamma

    def nextSeed(self):
        out = C.c_uint64()
        check(lib().cgb_randstate_next_seed(self._h, C.byref(out)))
        return out.value

    def getState(self):
        """the xoroshiro128+ state `Archive << randState` writes (math/Random.cpp:250-254,347-351)"""
        st = (C.c_uint64 * 2)()
        check(lib().cgb_randstate_get_state(self._h, st))
        return int(st[0]), int(st[1])

    def setState(self, state):
        st = (C.c_uint64 * 2)(int(state[0]), int(state[1]))
        check(lib().cgb_randstate_set_state(self._h, st))

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cgb_randstate_destroy(self._h)
            self._h = None


class GapsRng(object):
    """math/Random.cpp:32-200 (host stream)"""

    def __init__(self, randState):
        self._rs = randState
        self._h = C.c_void_p()
        check(lib().cgb_rng_create(randState._h, C.byref(self._h)))

    def getState(self):
        """the PCG state `Archive << rng` writes (math/Random.cpp:202-206)"""
        out = C.c_uint64()
        check(lib().cgb_rng_get_state(self._h, C.byref(out)))
        return out.value

    def setState(self, state):
        check(lib().cgb_rng_set_state(self._h, C.c_uint64(int(state))))

    def uniform32(self, a=None, b=None):
        out = C.c_uint32()
        if a is None:
            check(lib().cgb_rng_uniform32(self._h, C.byref(out)))
        else:
            check(lib().cgb_rng_uniform32_range(self._h, a, b, C.byref(out)))
        return out.value

    def uniform64(self, a, b):
        out = C.c_uint64()
        check(lib().cgb_rng_uniform64_range(self._h, a, b, C.byref(out)))
        return out.value

    def uniform(self):
        out = C.c_float()
        check(lib().cgb_rng_uniform(self._h, C.byref(out)))
        return out.value

    def poisson(self, lam):
        out = C.c_int32()
        check(lib().cgb_rng_poisson(self._h, lam, C.byref(out)))
        return out.value

    def exponential(self, lam):
        out = C.c_float()
        check(lib().cgb_rng_exponential(self._h, lam, C.byref(out)))
        return out.value

    def truncNormal(self, a, b, mean, sd):
        out, has = C.c_float(), C.c_int32()
        check(lib().cgb_rng_trunc_normal(self._h, a, b, mean, sd, C.byref(out), C.byref(has)))
        return out.value if has.value else None

    def truncGammaUpper(self, b, scale):
        out = C.c_float()
        check(lib().cgb_rng_trunc_gamma_upper(self._h, b, scale, C.byref(out)))
        return out.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cgb_rng_destroy(self._h)
            self._h = None


class GibbsSampler(object):
    """AsynchronousGibbsSampler<DenseNormalModel> on the device
    (gibbs_sampler/AsynchronousGibbsSampler.h:31-56, DenseNormalModel.h:15-64)."""

    def __init__(self, data, transpose, subsetRows, alpha, maxGibbsMass, params, randState):
        data = _f32(data)
        self._params = params if not isinstance(params, dict) else make_params(**params)
        self._rs = randState
        self._h = C.c_void_p()
        check(lib().cgb_sampler_create(fptr(data), data.shape[0], data.shape[1], 0, int(bool(transpose)),
                                       int(bool(subsetRows)), alpha, maxGibbsMass, C.byref(self._params),
                                       randState._h, C.byref(self._h)))
        rows, k, length = C.c_uint32(), C.c_uint32(), C.c_uint32()
        check(lib().cgb_sampler_shape(self._h, C.byref(rows), C.byref(k), C.byref(length)))
        self.nRows, self.nPatterns, self.rowLength = rows.value, k.value, length.value
        self._other = None

    def setUncertainty(self, unc, transpose, subsetRows, params=None):
        unc = _f32(unc)
        p = self._params if params is None else params
        check(lib().cgb_sampler_set_uncertainty(self._h, fptr(unc), unc.shape[0], unc.shape[1], 0,
                                                int(bool(transpose)), int(bool(subsetRows)), C.byref(p)))

    def setMatrix(self, mat):
        mat = _f32(mat)
        assert mat.shape == (self.nRows, self.nPatterns)
        check(lib().cgb_sampler_set_matrix(self._h, fptr(mat)))

    def setAnnealingTemp(self, temp):
        check(lib().cgb_sampler_set_annealing_temp(self._h, temp))

    def sync(self, other, nThreads=1):
        check(lib().cgb_sampler_sync(self._h, other._h))
        self._other = other

    def extraInitialization(self):
        check(lib().cgb_sampler_extra_initialization(self._h))

    def update(self, nSteps, nThreads=1):
        check(lib().cgb_sampler_update(self._h, nSteps, nThreads))

    def chiSq(self):
        out = C.c_float()
        check(lib().cgb_sampler_chisq(self._h, C.byref(out)))
        return out.value

    def nAtoms(self):
        out = C.c_uint64()
        check(lib().cgb_sampler_n_atoms(self._h, C.byref(out)))
        return out.value

    def dataSparsity(self):
        out = C.c_float()
        check(lib().cgb_sampler_data_sparsity(self._h, C.byref(out)))
        return out.value

    def getAverageQueueLength(self):
        out = C.c_float()
        check(lib().cgb_sampler_average_queue_length(self._h, C.byref(out)))
        return out.value

    def getMatrix(self):
        out = np.zeros((self.nRows, self.nPatterns), np.float32)
        check(lib().cgb_sampler_get_matrix(self._h, fptr(out)))
        return out

    def lambda_(self):
        lam, mx = C.c_float(), C.c_float()
        check(lib().cgb_sampler_lambda(self._h, C.byref(lam), C.byref(mx)))
        return lam.value, mx.value

    def atoms(self):
        n = C.c_uint64()
        check(lib().cgb_sampler_get_atoms(self._h, None, None, 0, C.byref(n)))
        pos = np.zeros(n.value, np.uint64)
        mass = np.zeros(n.value, np.float32)
        if n.value:
            check(lib().cgb_sampler_get_atoms(self._h, pos.ctypes.data_as(c_u64_p), fptr(mass), n.value, C.byref(n)))
        return pos, mass

    def setAtoms(self, pos, mass):
        """Replace the atomic domain; the order given becomes the pick order (ConcurrentAtomicDomain.cpp:144-155)."""
        pos = np.ascontiguousarray(pos, dtype=np.uint64)
        mass = _f32(mass)
        assert pos.shape == mass.shape
        check(lib().cgb_sampler_set_atoms(self._h, pos.ctypes.data_as(c_u64_p), fptr(mass), pos.size))

    def serialize(self):
        """`Archive << sampler` (AsynchronousGibbsSampler.h:221-226): bytes in the reference's wire format."""
        n = C.c_uint64()
        check(lib().cgb_sampler_serialize(self._h, None, 0, C.byref(n)))
        buf = (C.c_uint8 * n.value)()
        check(lib().cgb_sampler_serialize(self._h, buf, n.value, C.byref(n)))
        return bytes(buf)

    def deserialize(self, raw):
        """`Archive >> sampler` (:228-233); sync() / extraInitialization() are the caller's to repeat."""
        buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        check(lib().cgb_sampler_deserialize(self._h, buf, len(raw)))

    def apRow(self, row):
        out = np.zeros(self.rowLength, np.float32)
        check(lib().cgb_sampler_get_ap_row(self._h, row, fptr(out)))
        return out

    def alphaParameters(self, queries):
        """queries: rows of (variant, r1, c1, r2, c2, ch) — DenseNormalModel.cpp:162-240"""
        q = np.asarray(queries, dtype=np.float64).reshape(-1, 6)
        n = q.shape[0]
        variant = np.ascontiguousarray(q[:, 0], dtype=np.int32)
        r1, c1, r2, c2 = (np.ascontiguousarray(q[:, i], dtype=np.uint32) for i in (1, 2, 3, 4))
        ch = _f32(q[:, 5])
        s_out, smu_out = np.zeros(n, np.float32), np.zeros(n, np.float32)
        check(lib().cgb_sampler_alpha_parameters(
            self._h, n, variant.ctypes.data_as(c_i32_p), r1.ctypes.data_as(c_u32_p), c1.ctypes.data_as(c_u32_p),
            r2.ctypes.data_as(c_u32_p), c2.ctypes.data_as(c_u32_p), fptr(ch), fptr(s_out), fptr(smu_out)))
        return s_out, smu_out

    def counters(self):
        c = CgbSamplerCounters()
        check(lib().cgb_sampler_get_counters(self._h, C.byref(c)))
        return c

    def resetCounters(self):
        check(lib().cgb_sampler_reset_counters(self._h))

    def setKernelTiming(self, enabled):
        check(lib().cgb_sampler_set_kernel_timing(self._h, int(bool(enabled))))

    def setPersistent(self, enabled):
        check(lib().cgb_sampler_set_persistent(self._h, int(bool(enabled))))

    def setResidentShare(self, parts):
        """this sampler's resident grid takes 1/parts of the device (before its first update)"""
        check(lib().cgb_sampler_set_resident_share(self._h, int(parts)))

    def setUpdateMode(self, mode):
        """0: the reference's chain, proposal for proposal; 1: row-parallel sweep (cgb_sampler_set_update_mode)"""
        check(lib().cgb_sampler_set_update_mode(self._h, int(mode)))

    def reductionOrder(self):
        o = CgbReductionOrder()
        check(lib().cgb_sampler_reduction_order(self._h, C.byref(o)))
        return (o.threadsPerSegment, o.vectorWidth, o.nSegments, o.segmentLength)

    def deviceMatrix(self):
        dev, ld = C.c_void_p(), C.c_uint64()
        check(lib().cgb_sampler_device_matrix(self._h, C.byref(dev), C.byref(ld)))
        return dev.value, ld.value

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cgb_sampler_destroy(self._h)
            self._h = None


class GapsStatistics(object):
    """GapsStatistics.h:17-64, sums resident on the device"""

    def __init__(self, nGenes, nSamples, nPatterns):
        self.shape = (nGenes, nSamples, nPatterns)
        self._h = C.c_void_p()
        check(lib().cgb_stats_create(nGenes, nSamples, nPatterns, C.byref(self._h)))

    def update(self, A, P):
        check(lib().cgb_stats_update(self._h, A._h, P._h))

    def updateA(self, A, P):
        check(lib().cgb_stats_update_a(self._h, A._h, P._h))

    def updateP(self, A, P):
        check(lib().cgb_stats_update_p(self._h, A._h, P._h))

    def updatePump(self, A):
        check(lib().cgb_stats_update_pump(self._h, A._h))

    def _get(self, fn, rows):
        out = np.zeros((rows, self.shape[2]), np.float32)
        check(fn(self._h, fptr(out)))
        return out

    def Amean(self):
        return self._get(lib().cgb_stats_amean, self.shape[0])

    def Asd(self):
        return self._get(lib().cgb_stats_asd, self.shape[0])

    def Pmean(self):
        return self._get(lib().cgb_stats_pmean, self.shape[1])

    def Psd(self):
        return self._get(lib().cgb_stats_psd, self.shape[1])

    def pumpMatrix(self):
        return self._get(lib().cgb_stats_pump_matrix, self.shape[0])

    def meanPattern(self):
        return self._get(lib().cgb_stats_mean_pattern, self.shape[0])

    def meanChiSq(self, P):
        out = C.c_float()
        check(lib().cgb_stats_mean_chisq(self._h, P._h, C.byref(out)))
        return out.value

    def serialize(self):
        """`Archive << stats` (GapsStatistics.cpp:164-169)"""
        n = C.c_uint64()
        check(lib().cgb_stats_serialize(self._h, None, 0, C.byref(n)))
        buf = (C.c_uint8 * n.value)()
        check(lib().cgb_stats_serialize(self._h, buf, n.value, C.byref(n)))
        return bytes(buf)

    def deserialize(self, raw):
        buf = (C.c_uint8 * len(raw)).from_buffer_copy(raw)
        check(lib().cgb_stats_deserialize(self._h, buf, len(raw)))

    def deviceSums(self):
        a, a2, p, p2 = C.c_void_p(), C.c_void_p(), C.c_void_p(), C.c_void_p()
        lda, ldp, n = C.c_uint64(), C.c_uint64(), C.c_uint32()
        check(lib().cgb_stats_device_sums(self._h, C.byref(a), C.byref(a2), C.byref(p), C.byref(p2),
                                          C.byref(lda), C.byref(ldp), C.byref(n)))
        return dict(Amean=a.value, Asq=a2.value, Pmean=p.value, Psq=p2.value, ldA=lda.value, ldP=ldp.value,
                    nUpdates=n.value)

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cgb_stats_destroy(self._h)
            self._h = None


def reduction_order_for_length(rowLength):
    """(threadsPerSegment, vectorWidth, nSegments, segmentLength) of the eval kernel for rows of this length."""
    o = CgbReductionOrder()
    check(lib().cgb_reduction_order_for_length(rowLength, C.byref(o)))
    return (o.threadsPerSegment, o.vectorWidth, o.nSegments, o.segmentLength)


def sweep_reduction_order_for_length(rowLength):
    """The same for the row-parallel sweep (cgb_params.updateMode = UPDATE_SWEEP): one segment per row."""
    o = CgbReductionOrder()
    check(lib().cgb_sweep_reduction_order_for_length(rowLength, C.byref(o)))
    return (o.threadsPerSegment, o.vectorWidth, o.nSegments, o.segmentLength)


class Comm(object):
    """The C ABI's NCCL communicator (cgb_comm_*): the all-gather of per-shard factor rows straight from device memory
    (stitchTogether, R/DistributedCogaps.R:251-272) for callers without torch.  Rank 0 makes the id with
    Comm.unique_id(); every rank then builds Comm(id, rank, nRanks) on the device chosen with cgb_set_device."""

    @staticmethod
    def unique_id():
        buf = (C.c_uint8 * 128)()
        check(lib().cgb_comm_get_unique_id(buf))
        return bytes(buf)

    def __init__(self, unique_id, rank, nRanks):
        self._h = C.c_void_p()
        self.rank, self.nRanks = int(rank), int(nRanks)
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(lib().cgb_comm_init(buf, self.rank, self.nRanks, C.byref(self._h)))
        self.last_ms = 0.0

    def allgatherRows(self, sampler, rowsPerRank):
        """every rank's factor matrix stacked in rank order: (sum(rowsPerRank), nPatterns) float32"""
        rows = np.ascontiguousarray(rowsPerRank, dtype=np.uint32)
        assert rows.size == self.nRanks
        out = np.zeros((int(rows.sum()), sampler.nPatterns), np.float32)
        ms = C.c_double()
        check(lib().cgb_allgather_rows(self._h, sampler._h, rows.ctypes.data_as(c_u32_p), fptr(out), C.byref(ms)))
        self.last_ms = ms.value
        return out

    def allgatherDeviceRows(self, dev, ld, nPatterns, rowsPerRank):
        rows = np.ascontiguousarray(rowsPerRank, dtype=np.uint32)
        out = np.zeros((int(rows.sum()), nPatterns), np.float32)
        ms = C.c_double()
        check(lib().cgb_allgather_device_rows(self._h, C.c_void_p(dev), ld, nPatterns, rows.ctypes.data_as(c_u32_p), fptr(out), C.byref(ms)))
        self.last_ms = ms.value
        return out

    def __del__(self):
        if getattr(self, "_h", None):
            lib().cgb_comm_destroy(self._h)
            self._h = None
