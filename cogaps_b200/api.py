"""The user-facing surface of the reference, kept name for name so a CoGAPS user can switch:

    R: CoGAPS(data, params, nPatterns, nThreads, messages, outputFrequency, uncertainty, ..., ...)   R/CoGAPS.R:90-156
    R: new("CogapsParams", ...), setParam, getParam                                                 R/class-CogapsParams.R:44-193
    R: CogapsResult (featureLoadings, sampleFactors, loadingStdDev, factorStdDev, metadata)         R/class-CogapsResult.R:9-51
    C++: gaps::run(data, params, uncertainty, randState)                                            src/GapsRunner.h:14-24

R itself is not installed in this image, so this Python layer stands where R/CoGAPS.R stands: it validates
and packs parameters exactly like getGapsParameters (src/Cogaps.cpp:63-139) and calls the C ABI's cgb_run.
All computation happens in libcogaps_b200.so on the GPU.
"""
import ctypes as C
import copy
import os
import time

import numpy as np

from ._abi import PHASE_ALL, PHASE_EQUILIBRATION, PHASE_SAMPLING, CgbRunOptions, CgbCheckpointInfo, INTERRUPT_FN
from ._lib import lib, check
from ._runhelp import make_params, ResultArrays, fptr

_PARAM_DEFAULTS = dict(
    nPatterns=None, nIterations=50000, alphaA=0.01, alphaP=0.01, maxGibbsMassA=100.0, maxGibbsMassP=100.0,
    seed=None, sparseOptimization=False, distributed=None, nSets=4, cut=None, minNS=None, maxNS=None,
    explicitSets=None, samplingAnnotation=None, samplingWeight=None, subsetIndices=None, subsetDim=0,
    geneNames=None, sampleNames=None, checkpointInterval=0, checkpointInFile=None, checkpointOutFile=None,
    fixedPatterns=None, whichMatrixFixed="N", takePumpSamples=False,
)


class CogapsParams(object):
    """S4 class CogapsParams (R/class-CogapsParams.R:44-123): same slots, same defaults, same validity rules."""

    def __init__(self, nPatterns=7, distributed=None, **kw):
        for bad in ("nSets", "cut", "minNS", "maxNS"):
            if bad in kw:
                raise ValueError("%s must be set after CogapsParams are intialized" % bad)
        self.__dict__.update(_PARAM_DEFAULTS)
        if distributed == "none":
            distributed = None
        self.distributed = distributed
        self.nPatterns = nPatterns
        self.seed = int((time.time() % 1) * 1000) or 1
        self.cut = nPatterns
        self.minNS = int(np.ceil(self.nSets / 2.0))
        self.maxNS = self.minNS + self.nSets
        for k, v in kw.items():
            if k not in _PARAM_DEFAULTS:
                raise ValueError("invalid slot name %r for CogapsParams" % k)
            setattr(self, k, v)
        self.validate()

    def validate(self):
        """setValidity("CogapsParams"), R/class-CogapsParams.R:126-193"""
        def is_int(x):
            return float(x) == int(x)
        if self.nPatterns <= 0 or not is_int(self.nPatterns):
            raise ValueError("number of patterns must be an integer greater than zero")
        if self.nIterations <= 0 or not is_int(self.nIterations):
            raise ValueError("number of iterations must be an integer greater than zero")
        if self.alphaA <= 0 or self.alphaP <= 0:
            raise ValueError("alpha parameter must be greater than zero")
        if self.maxGibbsMassA <= 0 or self.maxGibbsMassP <= 0:
            raise ValueError("maxGibbsMass must be greater than zero")
        if self.seed <= 0 or not is_int(self.seed):
            raise ValueError("random seed must be an integer greater than zero")
        if self.whichMatrixFixed not in ("A", "P", "N"):
            raise ValueError("Invalid choice of fixed matrix, must be 'A' or 'P'")
        if self.fixedPatterns is not None and self.whichMatrixFixed == "N":
            raise ValueError("fixedPatterns passed without setting whichMatrixFixed")
        if self.whichMatrixFixed in ("A", "P") and self.fixedPatterns is None:
            raise ValueError("whichMatrixFixed is set without passing fixedPatterns")
        if self.subsetDim not in (0, 1, 2):
            raise ValueError("invalid subset dimension")
        if self.subsetDim > 0 and self.subsetIndices is None:
            raise ValueError("subsetDim provided without subsetIndices")
        if self.distributed is not None:
            if self.distributed not in ("genome-wide", "single-cell"):
                raise ValueError("distributed method must be either 'genome-wide' or 'single-cell'")
            if self.distributed == "single-cell" and self.whichMatrixFixed == "P":
                raise ValueError("can't fix P matrix when running single-cell CoGAPS")
            if self.distributed == "genome-wide" and self.whichMatrixFixed == "A":
                raise ValueError("can't fix A matrix when running genome-wide CoGAPS")
            if self.fixedPatterns is not None and self.explicitSets is None:
                raise ValueError("doing manual pattern matching without using explicit subsets")
            nGroups = len(set(self.samplingAnnotation)) if self.samplingAnnotation is not None else 0
            nWeights = len(self.samplingWeight) if self.samplingWeight is not None else 0
            if nGroups != nWeights:
                raise ValueError("samplingWeight has mismatched size with amount of distinct annotations")
            if self.cut > self.nPatterns:
                raise ValueError("cut must be less than or equal to nPatterns")
            if nWeights and not hasattr(self.samplingWeight, "keys"):
                raise ValueError("samplingWeight must be a named vector")
            if self.explicitSets is not None:
                if not isinstance(self.explicitSets, (list, tuple)):
                    raise ValueError("explicitSets must be a list")
                if len(self.explicitSets) != self.nSets:
                    raise ValueError("nSets doesn't match length of explicitSets")
                if self.samplingAnnotation is not None:
                    raise ValueError("explicitSets and samplingAnnotation/samplingWeight are both set")
                isChar = [all(isinstance(x, str) for x in s) for s in self.explicitSets]
                isNum = [all(not isinstance(x, str) for x in s) for s in self.explicitSets]
                if not all(isNum) and not all(isChar):
                    raise ValueError("explicitSets must be a list of numeric or character")
        return True

    def setParam(self, name, value):
        """setParam, R/methods-CogapsParams.R — nSets/nPatterns re-derive their dependents"""
        if name not in _PARAM_DEFAULTS:
            raise ValueError("invalid slot name %r for CogapsParams" % name)
        if name in ("samplingAnnotation", "samplingWeight"):
            raise ValueError("please set '%s' with setAnnotationWeights" % name)       # methods-CogapsParams.R:111-114
        before = dict(self.__dict__)                  # R objects have value semantics: a rejected change leaves no trace
        setattr(self, name, value)
        if name == "nSets":
            self.minNS = int(np.ceil(value / 2.0))
            self.maxNS = self.minNS + value
        if name == "nPatterns":
            self.cut = min(self.cut, value)
        try:
            self.validate()
        except ValueError:
            self.__dict__.clear()
            self.__dict__.update(before)
            raise
        return self

    def getParam(self, name):
        return getattr(self, name)

    def setAnnotationWeights(self, annotation, weights):
        """setAnnotationWeights (R/methods-CogapsParams.R:170-187): a label per row (column) being partitioned and a
        {label: weight} mapping for the weighted subsets of distributed CoGAPS"""
        before = (self.samplingAnnotation, self.samplingWeight)
        self.samplingAnnotation = list(annotation)
        self.samplingWeight = dict(weights)
        try:
            self.validate()
        except ValueError:
            self.samplingAnnotation, self.samplingWeight = before
            raise
        return self

    def setFixedPatterns(self, fixedPatterns, whichMatrixFixed):
        self.fixedPatterns = np.asarray(fixedPatterns, dtype=np.float32)
        self.whichMatrixFixed = whichMatrixFixed
        self.validate()
        return self


class CogapsResult(object):
    """CogapsResult (R/class-CogapsResult.R:9-51, built by createCogapsResult, R/methods-CogapsResult.R:8-51)"""

    def __init__(self, res, params, extras):
        self.featureLoadings = res.Amean      # Amean, nGenes x nPatterns
        self.sampleFactors = res.Pmean        # Pmean, nSamples x nPatterns
        self.loadingStdDev = res.Asd
        self.factorStdDev = res.Psd
        self.metadata = dict(
            meanChiSq=float(res.meanChiSq), chisq=res.chisqHistory, atomsA=res.atomHistoryA, atomsP=res.atomHistoryP,
            totalUpdates=int(res.totalUpdates), totalRunningTime=float(res.totalRunningTime),
            averageQueueLengthA=float(res.averageQueueLengthA), averageQueueLengthP=float(res.averageQueueLengthP),
            seed=int(res.seed), params=params, firstPassResults=None, unmatchedPatterns=None, clusteredPatterns=None,
            CorrToMeanPattern=None, subsets=None, version="cogaps_b200 (B200 device path)",
            pumpStat=res.pumpMatrix if params.takePumpSamples else None,
            meanPatternAssignment=res.meanPatternAssignment if params.takePumpSamples else None,
            equilibrationSnapshotsA=res.snapshotsA[:res.nSnapshotsEquilibration],
            equilibrationSnapshotsP=res.snapshotsP[:res.nSnapshotsEquilibration],
            samplingSnapshotsA=res.snapshotsA[res.nSnapshotsEquilibration:],
            samplingSnapshotsP=res.snapshotsP[res.nSnapshotsEquilibration:],
        )
        self.metadata.update(extras)

    # accessors named like the R generics
    def getFeatureLoadings(self):
        return self.featureLoadings

    def getAmplitudeMatrix(self):
        return self.featureLoadings

    def getSampleFactors(self):
        return self.sampleFactors

    def getPatternMatrix(self):
        return self.sampleFactors

    def getMeanChiSq(self):
        return self.metadata["meanChiSq"]


def _path(p):
    return None if p is None else os.fspath(p).encode()


def gaps_run(data, uncertainty=None, snapshots=False, checkpointInterval=0, checkpointOutFile=None,
             checkpointInFile=None, interrupt=None, **kw):
    """gaps::run (src/GapsRunner.h:14-24) through the C ABI: data nrow x ncol fp32 host array -> result arrays.
    Keyword arguments are the fields of cgb_params / GapsParameters.  The checkpoint arguments are the reference's
    (GapsParameters.h:37-38,46,56; GapsRunner.cpp:224-270): a checkpoint file is written every `checkpointInterval`
    iterations, `checkpointInFile` resumes from one (seed, nIterations, alphas, maxGibbsMass, sparseOptimization and
    the interval then come from the file).  `interrupt` is a callable polled once per iteration on this thread
    (Rcpp::checkUserInterrupt, GapsRunner.cpp:280); a true return stops the run with CogapsError(CGB_EINTERRUPTED)."""
    data = np.ascontiguousarray(data, dtype=np.float32)
    unc = np.ascontiguousarray(uncertainty, dtype=np.float32) if uncertainty is not None else None
    if unc is not None and unc.shape != data.shape:
        raise ValueError("uncertainty must have the same dimensions as the data")
    p = make_params(**kw)
    res = ResultArrays(p, data.shape[0], data.shape[1], snapshots=snapshots)
    if not checkpointInterval and checkpointInFile is None and interrupt is None:
        check(lib().cgb_run(fptr(data), data.shape[0], data.shape[1], 0, fptr(unc), C.byref(p), C.byref(res.c)))
        return res.finish()
    opt = CgbRunOptions()
    opt.struct_size = C.sizeof(CgbRunOptions)
    opt.checkpointInterval = int(checkpointInterval)
    opt.checkpointOutFile = _path(checkpointOutFile)
    opt.checkpointInFile = _path(checkpointInFile)
    callback = INTERRUPT_FN(lambda _user: 1 if interrupt() else 0) if interrupt is not None else None
    if callback is not None:
        opt.interrupt = callback        # `callback` stays referenced until the call returns
    check(lib().cgb_run_ex(fptr(data), data.shape[0], data.shape[1], 0, fptr(unc), C.byref(p), C.byref(opt), C.byref(res.c)))
    return res.finish()


def checkpoint_info(path):
    """What a checkpoint file holds (cgb_checkpoint_info_read): dict of the archived parameters, the phase and
    iteration a resumed run starts with, atom counts.  Host only."""
    info = CgbCheckpointInfo()
    info.struct_size = C.sizeof(CgbCheckpointInfo)
    check(lib().cgb_checkpoint_info_read(_path(path), C.byref(info)))
    return {name: getattr(info, name) for name, _ in CgbCheckpointInfo._fields_ if name != "struct_size"}


def checkpoint_rewrite(in_path, out_path):
    """Parse a checkpoint completely and write it back through the library's own writer (byte-identical for a file the
    reference wrote).  Host only."""
    check(lib().cgb_checkpoint_rewrite(_path(in_path), _path(out_path)))


def buildReport():
    """buildReport() (R/CoGAPS.R; getBuildReport_cpp, src/Cogaps.cpp:217-220): how the library was built"""
    return lib().cgb_build_report().decode()


def checkpointsEnabled():
    """checkpointsEnabled() (checkpointsEnabled_cpp, src/Cogaps.cpp:223-231): always — cgb_run_ex reads and writes the
    reference's checkpoint files"""
    return True


def compiledWithOpenMPSupport():
    """compiledWithOpenMPSupport() (src/Cogaps.cpp:234-242): no host thread team here; the fan-out over the proposals of a
    batch is the GPU"""
    return False


def getFileInfo(path):
    """getFileInfo_cpp (src/Cogaps.cpp:245-256): dimensions, rowNames, colNames of a data file.  As in the reference,
    colNames come from the header of .csv / .tsv files and rowNames are always empty.  Host only."""
    raw = os.fspath(path).encode()
    nrow, ncol = C.c_uint32(), C.c_uint32()
    check(lib().cgb_read_matrix_file(raw, None, 0, C.byref(nrow), C.byref(ncol)))
    needed, count = C.c_uint64(), C.c_uint32()
    check(lib().cgb_file_col_names(raw, None, 0, C.byref(needed), C.byref(count)))
    buf = C.create_string_buffer(max(int(needed.value), 1))
    check(lib().cgb_file_col_names(raw, buf, needed.value, C.byref(needed), C.byref(count)))
    names = buf.raw[:needed.value].split(b"\0")[:count.value]
    return {"dimensions": (nrow.value, ncol.value), "rowNames": [], "colNames": [n.decode(errors="replace") for n in names]}


def read_matrix_file(path):
    """A data / uncertainty file as the path overload of gaps::run reads it (.mtx, .csv, .tsv, .gct;
    src/file_parser/): nrow x ncol fp32 array.  Host only."""
    nrow, ncol = C.c_uint32(), C.c_uint32()
    check(lib().cgb_read_matrix_file(str(path).encode(), None, 0, C.byref(nrow), C.byref(ncol)))
    out = np.zeros((nrow.value, ncol.value), np.float32)
    check(lib().cgb_read_matrix_file(str(path).encode(), fptr(out), out.size, C.byref(nrow), C.byref(ncol)))
    return out


def write_matrix_csv(path, mat):
    """FileParser::writeToCsv (file_parser/FileParser.h:59-89): the matrix as the reference writes it.  Host only."""
    mat = np.ascontiguousarray(mat, dtype=np.float32)
    check(lib().cgb_write_matrix_csv(os.fspath(path).encode(), fptr(mat), mat.shape[0], mat.shape[1]))


def write_result_files(prefix, result):
    """GapsResult::writeToFile (GapsResult.cpp:27-35): <prefix>_<nPatterns>_{Amean,Pmean,Asd,Psd}.csv from a gaps_run
    result.  Host only."""
    check(lib().cgb_result_write_files(os.fspath(prefix).encode(), result.nGenes, result.nSamples, result.nPatterns,
                                       C.byref(result.c)))


def read_matrix_csr(path, by_rows=True):
    """The compressed rows (or columns) of a Matrix-Market file as the sparse model's loader builds them
    (cgb_read_matrix_csr): (nrow, ncol, ptr, idx, val).  Host only."""
    from ._abi import c_u32_p
    nrow, ncol, nnz = C.c_uint32(), C.c_uint32(), C.c_uint64()
    raw = os.fspath(path).encode()
    check(lib().cgb_read_matrix_csr(raw, int(bool(by_rows)), C.byref(nrow), C.byref(ncol), None, 0, None, None, 0, C.byref(nnz)))
    major = nrow.value if by_rows else ncol.value
    ptr = np.zeros(major + 1, np.uint32)
    idx = np.zeros(max(nnz.value, 1), np.uint32)
    val = np.zeros(max(nnz.value, 1), np.float32)
    check(lib().cgb_read_matrix_csr(raw, int(bool(by_rows)), C.byref(nrow), C.byref(ncol), ptr.ctypes.data_as(c_u32_p), ptr.size,
                                    idx.ctypes.data_as(c_u32_p), fptr(val), idx.size, C.byref(nnz)))
    return nrow.value, ncol.value, ptr, idx[:nnz.value], val[:nnz.value]


def gaps_run_file(path, uncertainty_path=None, snapshots=False, checkpointInterval=0, checkpointOutFile=None,
                  checkpointInFile=None, **kw):
    """gaps::run(const std::string &data, ...) (src/GapsRunner.h:19-24) through the C ABI; checkpoint arguments as in
    gaps_run."""
    nrow, ncol = C.c_uint32(), C.c_uint32()
    check(lib().cgb_read_matrix_file(str(path).encode(), None, 0, C.byref(nrow), C.byref(ncol)))
    p = make_params(**kw)
    res = ResultArrays(p, nrow.value, ncol.value, snapshots=snapshots)
    unc = str(uncertainty_path).encode() if uncertainty_path else None
    if not checkpointInterval and checkpointInFile is None:
        check(lib().cgb_run_file(str(path).encode(), unc, C.byref(p), C.byref(res.c)))
        return res.finish()
    opt = CgbRunOptions()
    opt.struct_size = C.sizeof(CgbRunOptions)
    opt.checkpointInterval = int(checkpointInterval)
    opt.checkpointOutFile = _path(checkpointOutFile)
    opt.checkpointInFile = _path(checkpointInFile)
    check(lib().cgb_run_file_ex(str(path).encode(), unc, C.byref(p), C.byref(opt), C.byref(res.c)))
    return res.finish()


_SNAPSHOT_PHASE = {"equilibration": PHASE_EQUILIBRATION, "sampling": PHASE_SAMPLING, "all": PHASE_ALL}


def CoGAPS(data, params=None, nPatterns=None, nThreads=1, messages=True, outputFrequency=1000, uncertainty=None,
           checkpointOutFile="gaps_checkpoint.out", checkpointInterval=0, checkpointInFile=None,
           transposeData=False, BPPARAM=None, workerID=1, asynchronousUpdates=True, nSnapshots=0,
           snapshotPhase="sampling", **kwargs):
    """CoGAPS() — R/CoGAPS.R:90-156.  `data` is a genes x samples array (or samples x genes with
    transposeData=True).  Extra keyword arguments overwrite slots of `params` (parseExtraParams,
    R/HelperFunctions.R:165-183: unknown names are an error)."""
    if params is None:
        params = CogapsParams(nPatterns=nPatterns if nPatterns is not None else 7)
    else:
        params = copy.copy(params)          # R passes parameters by value: the caller's object is never changed
        if nPatterns is not None:
            params.setParam("nPatterns", nPatterns)
    for k, v in kwargs.items():
        if k not in _PARAM_DEFAULTS:
            raise ValueError("unrecognized argument: %s" % k)
        params.setParam(k, v)
    params.validate()
    # `data` may be the name of a .mtx / .csv / .tsv / .gct file (cogaps_from_file_cpp, src/Cogaps.cpp:188-202): the
    # library reads it itself; the matrix checks below are for in-memory data only, as in R (checkDataMatrix)
    fromFile = isinstance(data, (str, os.PathLike))
    if fromFile:
        if uncertainty is not None and not isinstance(uncertainty, (str, os.PathLike)):
            raise ValueError("uncertainty must be same data type as data (file name)")   # R/HelperFunctions.R:207-208
        fileShape = getFileInfo(data)["dimensions"]      # also: "unsupported file extension" (R/HelperFunctions.R:8-12)
        if uncertainty is not None and params.sparseOptimization:
            raise ValueError("must use default uncertainty when enabling sparseOptimization")
    else:
        if isinstance(uncertainty, (str, os.PathLike)):
            raise ValueError("uncertainty must be a matrix unless data is a file path")  # R/HelperFunctions.R:209-210
        data = np.asarray(data)
    # checkInputs, R/HelperFunctions.R:203-260
    if not fromFile and data.ndim != 2:
        raise ValueError("data must be a matrix")
    if not fromFile and np.isnan(data).any():
        raise ValueError("NA values in data")
    if not fromFile and (data < 0).any():
        raise ValueError("negative values in data matrix")
    if uncertainty is not None and not fromFile:
        uncertainty = np.asarray(uncertainty)
        if (uncertainty < 0).any():
            raise ValueError("negative values in uncertainty matrix")
        if params.sparseOptimization:
            raise ValueError("must use default uncertainty when enabling sparseOptimization")
    checkpointing = checkpointInFile is not None or bool(checkpointInterval)
    if checkpointing and params.distributed is not None:
        raise ValueError("checkpoints not supported for distributed cogaps")      # R/HelperFunctions.R:221-222
    if checkpointing and params.subsetDim:
        checkpointInterval = 0                    # createCheckpoint skips subset runs (GapsRunner.cpp:232)
    if checkpointing and not asynchronousUpdates:
        raise ValueError("checkpoints need asynchronousUpdates=True: the reference's sequential sampler cannot be "
                         "resumed either (SingleThreadedGibbsSampler.h:260-273)")
    if params.distributed is not None:
        from .distributed import distributedCogaps
        # the reference forces the sequential sampler here (R/DistributedCogaps.R:28-29); we follow the caller
        return distributedCogaps(data, params, uncertainty, nThreads=nThreads, messages=messages,
                                 outputFrequency=outputFrequency, transposeData=transposeData,
                                 sequentialSampler=not asynchronousUpdates)
    shape = fileShape if fromFile else data.shape
    nGenes, nSamples = (shape[1], shape[0]) if transposeData else shape
    if params.nPatterns >= min(nGenes, nSamples) and params.subsetDim == 0:
        pass  # R only warns here
    kw = dict(seed=int(params.seed), nPatterns=int(params.nPatterns), nIterations=int(params.nIterations),
              maxThreads=int(nThreads), outputFrequency=int(outputFrequency), alphaA=params.alphaA,
              alphaP=params.alphaP, maxGibbsMassA=params.maxGibbsMassA, maxGibbsMassP=params.maxGibbsMassP,
              transposeData=int(bool(transposeData)), useSparseOptimization=int(bool(params.sparseOptimization)),
              asynchronousUpdates=int(bool(asynchronousUpdates)), takePumpSamples=int(bool(params.takePumpSamples)),
              printMessages=int(bool(messages) and workerID == 1), whichMatrixFixed=params.whichMatrixFixed,
              workerID=int(workerID))
    if checkpointInFile is not None:
        # the run continues with the archived nPatterns and nIterations (run_helper, GapsRunner.cpp:99-105); the result
        # arrays (histories, snapshots) must be sized for them
        info = checkpoint_info(checkpointInFile)
        kw["nPatterns"] = int(info["nPatterns"])
        kw["nIterations"] = int(info["nIterations"])
    if nSnapshots:
        # Cogaps.cpp:95-99: plain integer division — more snapshots asked for than iterations means none are taken
        kw["snapshotFrequency"] = int(kw["nIterations"]) // int(nSnapshots)
        kw["snapshotPhase"] = _SNAPSHOT_PHASE[snapshotPhase]
    if params.subsetDim:
        kw["subsetGenes"] = 1 if params.subsetDim == 1 else 0                          # Cogaps.cpp:120-126
        kw["subsetIndices"] = np.asarray(params.subsetIndices, dtype=np.uint32)
    if params.fixedPatterns is not None:
        kw["fixedPatterns"] = np.asarray(params.fixedPatterns, dtype=np.float32)
    if fromFile:
        res = gaps_run_file(data, uncertainty_path=uncertainty, snapshots=bool(kw.get("snapshotFrequency")),
                            checkpointInterval=int(checkpointInterval or 0), checkpointOutFile=checkpointOutFile,
                            checkpointInFile=checkpointInFile, **kw)
    else:
        res = gaps_run(data, uncertainty=uncertainty, snapshots=bool(kw.get("snapshotFrequency")), checkpointInterval=int(checkpointInterval or 0),
                       checkpointOutFile=checkpointOutFile, checkpointInFile=checkpointInFile, **kw)
    extras = dict(nBatchesA=res.nBatchesA, nBatchesP=res.nBatchesP, secondsUpdateA=res.secondsUpdateA,
                  secondsUpdateP=res.secondsUpdateP, algorithmicBytes=res.algorithmicBytes)
    return CogapsResult(res, params, extras)
