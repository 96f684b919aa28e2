"""ctypes loaders for the two CPU checkers.  TEST INFRASTRUCTURE ONLY.

  RefLib(variant)  -> oracle/_ref/libcogaps_ref_<variant>.so  (the unmodified reference, compiled in place)
  OracleLib()      -> oracle/libcogaps_oracle.so              (our C restatement)

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import this.
"""
import ctypes as C
import os
import numpy as np

from cogaps_b200._abi import (CgbParams, CgbResult, c_float_p, c_u32_p, c_u64_p, c_i32_p,
                              ERF_TABLE_SIZE, ERFINV_TABLE_SIZE, QGAMMA_TABLE_SIZE)
from cogaps_b200._runhelp import make_params, ResultArrays, fptr

HERE = os.path.dirname(os.path.abspath(__file__))


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _u32(a):
    return np.ascontiguousarray(a, dtype=np.uint32)


def _path(p):
    return None if p is None else os.fspath(p).encode()


class RefLib(object):
    """The reference C++ core behind oracle/ref_driver.cpp."""

    def __init__(self, variant="scalar"):
        path = os.path.join(HERE, "_ref", "libcogaps_ref_%s.so" % variant)
        if not os.path.exists(path):
            raise OSError("reference build %s missing - run `make -C oracle ref` where /root/reference exists" % path)
        self.variant = variant
        self.lib = C.CDLL(path)
        self.lib.cogaps_ref_build_report.restype = C.c_char_p
        self.lib.cogaps_ref_rng_stream.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_uint64, C.c_uint64,
                                                   C.c_double, C.c_float, C.c_float, C.c_float, C.c_float, c_u64_p]

    @staticmethod
    def available(variant="scalar"):
        return os.path.exists(os.path.join(HERE, "_ref", "libcogaps_ref_%s.so" % variant))

    def build_report(self):
        return self.lib.cogaps_ref_build_report().decode()

    def max_threads(self):
        return self.lib.cogaps_ref_max_threads()

    def alpha_parameters_sparse(self, data, A, P, queries):
        """SparseNormalModel::alphaParameters* of the reference itself; queries as in alpha_parameters."""
        data, A, P = _f32(data), _f32(A), _f32(P)
        g, s = data.shape
        k = A.shape[1]
        q = np.asarray(queries, dtype=np.float64).reshape(-1, 6)
        n = q.shape[0]
        variant = np.ascontiguousarray(q[:, 0], dtype=np.int32)
        r1, c1, r2, c2 = (_u32(q[:, i]) for i in (1, 2, 3, 4))
        ch = _f32(q[:, 5])
        s_out = np.zeros(n, np.float32)
        smu_out = np.zeros(n, np.float32)
        rc = self.lib.cogaps_ref_alpha_parameters_sparse(
            fptr(data), C.c_uint32(g), C.c_uint32(s), C.c_uint32(k), fptr(A), fptr(P),
            C.c_uint32(n), variant.ctypes.data_as(c_i32_p), r1.ctypes.data_as(c_u32_p),
            c1.ctypes.data_as(c_u32_p), r2.ctypes.data_as(c_u32_p), c2.ctypes.data_as(c_u32_p),
            fptr(ch), fptr(s_out), fptr(smu_out))
        if rc != 0:
            raise RuntimeError("cogaps_ref_alpha_parameters_sparse failed")
        return s_out, smu_out

    def chisq_sparse(self, data, A, P):
        data, A, P = _f32(data), _f32(A), _f32(P)
        out = np.zeros(3, np.float32)
        rc = self.lib.cogaps_ref_chisq_sparse(fptr(data), C.c_uint32(data.shape[0]), C.c_uint32(data.shape[1]),
                                              C.c_uint32(A.shape[1]), fptr(A), fptr(P), fptr(out))
        if rc != 0:
            raise RuntimeError("cogaps_ref_chisq_sparse failed")
        return out

    def read_file(self, path):
        """Matrix(path, ...) through the reference's own file parsers."""
        nrow, ncol = C.c_uint32(), C.c_uint32()
        self.lib.cogaps_ref_read_file(str(path).encode(), None, C.byref(nrow), C.byref(ncol))
        out = np.zeros((nrow.value, ncol.value), np.float32)
        self.lib.cogaps_ref_read_file(str(path).encode(), fptr(out), C.byref(nrow), C.byref(ncol))
        return out

    def file_info(self, path):
        """getFileInfo_cpp through the reference's own FileParser: (nrow, ncol), rowNames, colNames"""
        nrow, ncol = C.c_uint32(), C.c_uint32()
        counts = (C.c_uint32 * 2)()
        buf = C.create_string_buffer(1 << 20)
        rc = self.lib.cogaps_ref_file_info(_path(path), C.byref(nrow), C.byref(ncol), buf, C.c_uint64(len(buf)), counts)
        if rc != 0:
            raise RuntimeError("cogaps_ref_file_info failed")
        names = buf.raw.split(b"\0")[:counts[0] + counts[1]]
        names = [n.decode(errors="replace") for n in names]
        return (nrow.value, ncol.value), names[:counts[0]], names[counts[0]:]

    def write_csv(self, path, mat):
        """FileParser::writeToCsv of the reference itself."""
        mat = _f32(mat)
        self.lib.cogaps_ref_write_csv(_path(path), fptr(mat), C.c_uint32(mat.shape[0]), C.c_uint32(mat.shape[1]))

    def run(self, data, uncertainty=None, snapshots=False, checkpointInterval=0, checkpointOutFile=None,
            checkpointInFile=None, **kw):
        """gaps::run; the checkpoint arguments are the reference's own (GapsParameters.h:37-38,46,56)."""
        data = _f32(data)
        unc = _f32(uncertainty) if uncertainty is not None else None
        p = make_params(**kw)
        res = ResultArrays(p, data.shape[0], data.shape[1], snapshots=snapshots)
        rc = self.lib.cogaps_ref_run_checkpointed(
            fptr(data), C.c_uint32(data.shape[0]), C.c_uint32(data.shape[1]), fptr(unc), C.byref(p), C.byref(res.c),
            C.c_uint32(checkpointInterval), _path(checkpointOutFile), _path(checkpointInFile))
        if rc != 0:
            raise RuntimeError("cogaps_ref_run failed: %d" % rc)
        out = res.finish()
        # the reference's own phases (its clock readings at GapsRunner.cpp:400-473): seconds before the sampler loop
        # and inside it; totalRunningTime above is the wall time of the whole gaps::run call
        load, loop = C.c_double(), C.c_double()
        self.lib.cogaps_ref_last_run_seconds(C.byref(load), C.byref(loop))
        out.secondsLoading, out.secondsSamplerLoop = load.value, loop.value
        return out

    def tables(self):
        erf = np.zeros(ERF_TABLE_SIZE, np.float32)
        erfinv = np.zeros(ERFINV_TABLE_SIZE, np.float32)
        qgamma = np.zeros(QGAMMA_TABLE_SIZE, np.float32)
        self.lib.cogaps_ref_tables(fptr(erf), fptr(erfinv), fptr(qgamma))
        return erf, erfinv, qgamma

    def rng_stream(self, seed, kind, n, a=0, b=0, lam=0.0, f=(0.0, 0.0, 0.0, 0.0)):
        out = np.zeros(n, np.uint64)
        rc = self.lib.cogaps_ref_rng_stream(seed, kind, n, a, b, lam, f[0], f[1], f[2], f[3],
                                            out.ctypes.data_as(c_u64_p))
        if rc != 0:
            raise RuntimeError("cogaps_ref_rng_stream failed")
        return out

    def alpha_parameters(self, data, A, P, queries, uncertainty=None, want_ap=False):
        """queries: list of (variant, r1, c1, r2, c2, ch)."""
        data, A, P = _f32(data), _f32(A), _f32(P)
        g, s = data.shape
        k = A.shape[1]
        q = np.asarray(queries, dtype=np.float64).reshape(-1, 6)
        n = q.shape[0]
        variant = np.ascontiguousarray(q[:, 0], dtype=np.int32)
        r1, c1, r2, c2 = (_u32(q[:, i]) for i in (1, 2, 3, 4))
        ch = _f32(q[:, 5])
        s_out = np.zeros(n, np.float32)
        smu_out = np.zeros(n, np.float32)
        ap = np.zeros((g, s), np.float32) if want_ap else None
        unc = _f32(uncertainty) if uncertainty is not None else None
        rc = self.lib.cogaps_ref_alpha_parameters(
            fptr(data), C.c_uint32(g), C.c_uint32(s), C.c_uint32(k), fptr(A), fptr(P), fptr(unc),
            C.c_uint32(n), variant.ctypes.data_as(c_i32_p), r1.ctypes.data_as(c_u32_p),
            c1.ctypes.data_as(c_u32_p), r2.ctypes.data_as(c_u32_p), c2.ctypes.data_as(c_u32_p),
            fptr(ch), fptr(s_out), fptr(smu_out), fptr(ap))
        if rc != 0:
            raise RuntimeError("cogaps_ref_alpha_parameters failed")
        return (s_out, smu_out, ap) if want_ap else (s_out, smu_out)

    def chisq(self, data, A, P, uncertainty=None):
        data, A, P = _f32(data), _f32(A), _f32(P)
        unc = _f32(uncertainty) if uncertainty is not None else None
        out = np.zeros(3, np.float32)
        rc = self.lib.cogaps_ref_chisq(fptr(data), C.c_uint32(data.shape[0]), C.c_uint32(data.shape[1]),
                         C.c_uint32(A.shape[1]), fptr(A), fptr(P), fptr(unc), fptr(out))
        if rc != 0:
            raise RuntimeError("chisq probe failed")
        return out


class OracleOptions(C.Structure):
    _fields_ = [
        ("reduceMode", C.c_int32),
        ("mathMode", C.c_int32),
        ("orderA", C.c_uint32 * 4),
        ("orderP", C.c_uint32 * 4),
        ("erf", c_float_p),
        ("erfinv", c_float_p),
        ("qgamma", c_float_p),
        ("checkpointInterval", C.c_uint32),
        ("checkpointOutFile", C.c_char_p),
        ("checkpointInFile", C.c_char_p),
        ("stopAfterCheckpoints", C.c_uint32),
    ]


class TraceRecord(C.Structure):
    _fields_ = [
        ("phase", C.c_uint32), ("iter", C.c_uint32), ("side", C.c_uint32), ("batch", C.c_uint32),
        ("type", C.c_uint32), ("r1", C.c_uint32), ("c1", C.c_uint32), ("r2", C.c_uint32), ("c2", C.c_uint32),
        ("accepted", C.c_uint32), ("pos", C.c_uint64), ("atom1Pos", C.c_uint64), ("atom2Pos", C.c_uint64),
        ("rngState", C.c_uint64), ("mass1", C.c_float), ("mass2", C.c_float), ("newMass1", C.c_float),
        ("newMass2", C.c_float), ("s", C.c_float), ("s_mu", C.c_float),
    ]


TRACE_DTYPE = np.dtype([(n, np.dtype(t)) for n, t in TraceRecord._fields_], align=True)
REDUCE_SCALAR, REDUCE_AVX8, REDUCE_DEVICE = 0, 1, 2
MATH_LIBM, MATH_PORTABLE = 0, 1


class OracleLib(object):
    """Our C restatement (oracle/cogaps_oracle.c)."""

    def __init__(self):
        path = os.path.join(HERE, "libcogaps_oracle.so")
        if not os.path.exists(path):
            raise OSError("%s missing - run `make -C oracle oracle`" % path)
        self.lib = C.CDLL(path)
        self.lib.cogaps_oracle_rng_stream.argtypes = [C.c_uint32, C.c_int, C.c_uint32, C.c_uint64, C.c_uint64,
                                                      C.c_double, C.c_float, C.c_float, C.c_float, C.c_float, c_u64_p]
        self.lib.cogaps_oracle_portable_logf.restype = C.c_float
        self.lib.cogaps_oracle_portable_logf.argtypes = [C.c_float]

    @staticmethod
    def options(reduce="scalar", math="libm", orderA=None, orderP=None, tables=None, checkpointInterval=0,
                checkpointOutFile=None, checkpointInFile=None, stopAfterCheckpoints=0):
        o = OracleOptions()
        o.reduceMode = {"scalar": REDUCE_SCALAR, "avx8": REDUCE_AVX8, "device": REDUCE_DEVICE}[reduce]
        o.mathMode = {"libm": MATH_LIBM, "portable": MATH_PORTABLE}[math]
        if orderA is not None:
            o.orderA = (C.c_uint32 * 4)(*orderA)
        if orderP is not None:
            o.orderP = (C.c_uint32 * 4)(*orderP)
        keep = []
        if tables is not None:
            erf, erfinv, qgamma = (_f32(t) for t in tables)
            keep = [erf, erfinv, qgamma]
            o.erf, o.erfinv, o.qgamma = fptr(erf), fptr(erfinv), fptr(qgamma)
        o.checkpointInterval = checkpointInterval
        o.checkpointOutFile = _path(checkpointOutFile)
        o.checkpointInFile = _path(checkpointInFile)
        o.stopAfterCheckpoints = stopAfterCheckpoints
        o._keepalive = keep
        return o

    def run(self, data, uncertainty=None, snapshots=False, options=None, trace_capacity=0, **kw):
        data = _f32(data)
        unc = _f32(uncertainty) if uncertainty is not None else None
        p = make_params(**kw)
        res = ResultArrays(p, data.shape[0], data.shape[1], snapshots=snapshots)
        opt = options if options is not None else self.options()
        if trace_capacity:
            trace = np.zeros(trace_capacity, TRACE_DTYPE)
            count = C.c_uint64(0)
            rc = self.lib.cogaps_oracle_run_trace(
                fptr(data), C.c_uint32(data.shape[0]), C.c_uint32(data.shape[1]), fptr(unc), C.byref(p),
                C.byref(res.c), C.byref(opt), trace.ctypes.data_as(C.c_void_p), C.c_uint64(trace_capacity),
                C.byref(count))
        else:
            rc = self.lib.cogaps_oracle_run(fptr(data), C.c_uint32(data.shape[0]), C.c_uint32(data.shape[1]),
                                            fptr(unc), C.byref(p), C.byref(res.c), C.byref(opt))
        if rc == -7 and opt.stopAfterCheckpoints:
            return None     # stopped on request right after a checkpoint; the file is the product of the call
        if rc != 0:
            raise RuntimeError("cogaps_oracle_run failed: %d" % rc)
        out = res.finish()
        if trace_capacity:
            out.trace_total = count.value
            out.trace = trace[:min(count.value, trace_capacity)]
        return out

    def tables(self):
        erf = np.zeros(ERF_TABLE_SIZE, np.float32)
        erfinv = np.zeros(ERFINV_TABLE_SIZE, np.float32)
        qgamma = np.zeros(QGAMMA_TABLE_SIZE, np.float32)
        self.lib.cogaps_oracle_tables(fptr(erf), fptr(erfinv), fptr(qgamma))
        return erf, erfinv, qgamma

    def rng_stream(self, seed, kind, n, a=0, b=0, lam=0.0, f=(0.0, 0.0, 0.0, 0.0)):
        out = np.zeros(n, np.uint64)
        rc = self.lib.cogaps_oracle_rng_stream(seed, kind, n, a, b, lam, f[0], f[1], f[2], f[3],
                                               out.ctypes.data_as(c_u64_p))
        if rc != 0:
            raise RuntimeError("cogaps_oracle_rng_stream failed")
        return out

    def portable_logf(self, x):
        return self.lib.cogaps_oracle_portable_logf(float(x))

    def alpha_parameters_sparse(self, data, A, P, queries, options=None):
        """SparseNormalModel.cpp:153-292 through the oracle; queries as in alpha_parameters."""
        data, A, P = _f32(data), _f32(A), _f32(P)
        g, s = data.shape
        k = A.shape[1]
        q = np.asarray(queries, dtype=np.float64).reshape(-1, 6)
        n = q.shape[0]
        variant = np.ascontiguousarray(q[:, 0], dtype=np.int32)
        r1, c1, r2, c2 = (_u32(q[:, i]) for i in (1, 2, 3, 4))
        ch = _f32(q[:, 5])
        s_out = np.zeros(n, np.float32)
        smu_out = np.zeros(n, np.float32)
        opt = options if options is not None else self.options()
        rc = self.lib.cogaps_oracle_alpha_parameters_sparse(
            fptr(data), C.c_uint32(g), C.c_uint32(s), C.c_uint32(k), fptr(A), fptr(P),
            C.c_uint32(n), variant.ctypes.data_as(c_i32_p), r1.ctypes.data_as(c_u32_p),
            c1.ctypes.data_as(c_u32_p), r2.ctypes.data_as(c_u32_p), c2.ctypes.data_as(c_u32_p),
            fptr(ch), fptr(s_out), fptr(smu_out), C.byref(opt))
        if rc != 0:
            raise RuntimeError("cogaps_oracle_alpha_parameters_sparse failed")
        return s_out, smu_out

    def chisq_sparse(self, data, A, P):
        data, A, P = _f32(data), _f32(A), _f32(P)
        out = np.zeros(3, np.float32)
        rc = self.lib.cogaps_oracle_chisq_sparse(fptr(data), C.c_uint32(data.shape[0]), C.c_uint32(data.shape[1]),
                         C.c_uint32(A.shape[1]), fptr(A), fptr(P), fptr(out))
        if rc != 0:
            raise RuntimeError("sparse chisq probe failed")
        return out

    def alpha_parameters(self, data, A, P, queries, uncertainty=None, want_ap=False, options=None):
        data, A, P = _f32(data), _f32(A), _f32(P)
        g, s = data.shape
        k = A.shape[1]
        q = np.asarray(queries, dtype=np.float64).reshape(-1, 6)
        n = q.shape[0]
        variant = np.ascontiguousarray(q[:, 0], dtype=np.int32)
        r1, c1, r2, c2 = (_u32(q[:, i]) for i in (1, 2, 3, 4))
        ch = _f32(q[:, 5])
        s_out = np.zeros(n, np.float32)
        smu_out = np.zeros(n, np.float32)
        ap = np.zeros((g, s), np.float32) if want_ap else None
        unc = _f32(uncertainty) if uncertainty is not None else None
        opt = options if options is not None else self.options()
        rc = self.lib.cogaps_oracle_alpha_parameters(
            fptr(data), C.c_uint32(g), C.c_uint32(s), C.c_uint32(k), fptr(A), fptr(P), fptr(unc),
            C.c_uint32(n), variant.ctypes.data_as(c_i32_p), r1.ctypes.data_as(c_u32_p),
            c1.ctypes.data_as(c_u32_p), r2.ctypes.data_as(c_u32_p), c2.ctypes.data_as(c_u32_p),
            fptr(ch), fptr(s_out), fptr(smu_out), fptr(ap), C.byref(opt))
        if rc != 0:
            raise RuntimeError("cogaps_oracle_alpha_parameters failed")
        return (s_out, smu_out, ap) if want_ap else (s_out, smu_out)

    def chisq(self, data, A, P, uncertainty=None):
        data, A, P = _f32(data), _f32(A), _f32(P)
        unc = _f32(uncertainty) if uncertainty is not None else None
        out = np.zeros(3, np.float32)
        rc = self.lib.cogaps_oracle_chisq(fptr(data), C.c_uint32(data.shape[0]), C.c_uint32(data.shape[1]),
                         C.c_uint32(A.shape[1]), fptr(A), fptr(P), fptr(unc), fptr(out))
        if rc != 0:
            raise RuntimeError("chisq probe failed")
        return out
