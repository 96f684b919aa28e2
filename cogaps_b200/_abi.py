"""ctypes mirror of include/cogaps_b200.h (structs and constants only — no library is loaded here)."""
import ctypes as C

CGB_OK = 0
CGB_EINVAL = -1
CGB_ENODEVICE = -2
CGB_ECUDA = -3
CGB_ENOMEM = -4
CGB_EUNSUPPORTED = -5
CGB_EINTERNAL = -6
CGB_EINTERRUPTED = -7

ERF_TABLE_SIZE = 3001
ERFINV_TABLE_SIZE = 5001
QGAMMA_TABLE_SIZE = 5001

PHASE_EQUILIBRATION = 1
PHASE_SAMPLING = 2
PHASE_ALL = 3

c_float_p = C.POINTER(C.c_float)
c_u32_p = C.POINTER(C.c_uint32)
c_u64_p = C.POINTER(C.c_uint64)
c_i32_p = C.POINTER(C.c_int32)


class CgbParams(C.Structure):
    """struct cgb_params — mirrors GapsParameters (reference src/GapsParameters.h:25-70)."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("seed", C.c_uint32),
        ("nPatterns", C.c_uint32),
        ("nIterations", C.c_uint32),
        ("maxThreads", C.c_uint32),
        ("outputFrequency", C.c_uint32),
        ("snapshotFrequency", C.c_uint32),
        ("snapshotPhase", C.c_uint32),
        ("alphaA", C.c_float),
        ("alphaP", C.c_float),
        ("maxGibbsMassA", C.c_float),
        ("maxGibbsMassP", C.c_float),
        ("transposeData", C.c_int32),
        ("useSparseOptimization", C.c_int32),
        ("asynchronousUpdates", C.c_int32),
        ("takePumpSamples", C.c_int32),
        ("printMessages", C.c_int32),
        ("whichMatrixFixed", C.c_int32),
        ("subsetGenes", C.c_int32),
        ("nSubsetIndices", C.c_uint32),
        ("subsetIndices", c_u32_p),
        ("fixedPatterns", c_float_p),
        ("workerID", C.c_uint32),
        ("runningDistributed", C.c_int32),
        ("updateMode", C.c_int32),
    ]

    @classmethod
    def defaults(cls):
        """Reference defaults (src/GapsParameters.h:79-114)."""
        p = cls()
        p.struct_size = C.sizeof(cls)
        p.seed = 0
        p.nPatterns = 3
        p.nIterations = 1000
        p.maxThreads = 1
        p.outputFrequency = 500
        p.snapshotFrequency = 0
        p.snapshotPhase = PHASE_ALL
        p.alphaA = 0.01
        p.alphaP = 0.01
        p.maxGibbsMassA = 100.0
        p.maxGibbsMassP = 100.0
        p.transposeData = 0
        p.useSparseOptimization = 0
        p.asynchronousUpdates = 1
        p.takePumpSamples = 0
        p.printMessages = 0
        p.whichMatrixFixed = ord("N")
        p.subsetGenes = 0
        p.nSubsetIndices = 0
        p.workerID = 1
        p.runningDistributed = 0
        return p


class CgbResult(C.Structure):
    """struct cgb_result — mirrors GapsResult (reference src/GapsResult.h:11-36)."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("historyCapacity", C.c_uint32),
        ("Amean", c_float_p),
        ("Asd", c_float_p),
        ("Pmean", c_float_p),
        ("Psd", c_float_p),
        ("chisqHistory", c_float_p),
        ("atomHistoryA", c_u32_p),
        ("atomHistoryP", c_u32_p),
        ("pumpMatrix", c_float_p),
        ("meanPatternAssignment", c_float_p),
        ("snapshotsA", c_float_p),
        ("snapshotsP", c_float_p),
        ("snapshotCapacity", C.c_uint32),
        ("nSnapshotsEquilibration", C.c_uint32),
        ("nSnapshotsSampling", C.c_uint32),
        ("nHistory", C.c_uint32),
        ("seed", C.c_uint32),
        ("totalUpdates", C.c_uint64),
        ("totalRunningTime", C.c_double),
        ("meanChiSq", C.c_float),
        ("averageQueueLengthA", C.c_float),
        ("averageQueueLengthP", C.c_float),
        ("nBatchesA", C.c_uint64),
        ("nBatchesP", C.c_uint64),
        ("secondsUpdateA", C.c_double),
        ("secondsUpdateP", C.c_double),
        ("secondsDevice", C.c_double),
        ("algorithmicBytes", C.c_double),
    ]


class CgbSamplerCounters(C.Structure):
    _fields_ = [
        ("nBatches", C.c_uint64),
        ("nProposalsQueued", C.c_uint64),
        ("nProposalsTotal", C.c_uint64),
        ("algorithmicBytes", C.c_double),
        ("secondsHostGenerate", C.c_double),
        ("secondsDeviceWait", C.c_double),
        ("secondsKernel", C.c_double),
    ]


class CgbReductionOrder(C.Structure):
    _fields_ = [
        ("threadsPerSegment", C.c_uint32),
        ("vectorWidth", C.c_uint32),
        ("nSegments", C.c_uint32),
        ("segmentLength", C.c_uint32),
    ]


INTERRUPT_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p)


class CgbRunOptions(C.Structure):
    """struct cgb_run_options — checkpoints (GapsParameters.h:37-38,46,56) and the per-iteration interrupt poll."""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("checkpointInterval", C.c_uint32),
        ("checkpointOutFile", C.c_char_p),
        ("checkpointInFile", C.c_char_p),
        ("interrupt", INTERRUPT_FN),
        ("interruptUser", C.c_void_p),
    ]


class CgbCheckpointInfo(C.Structure):
    """struct cgb_checkpoint_info"""
    _fields_ = [
        ("struct_size", C.c_uint32),
        ("seed", C.c_uint32),
        ("nGenes", C.c_uint32),
        ("nSamples", C.c_uint32),
        ("nPatterns", C.c_uint32),
        ("nIterations", C.c_uint32),
        ("alphaA", C.c_float),
        ("alphaP", C.c_float),
        ("maxGibbsMassA", C.c_float),
        ("maxGibbsMassP", C.c_float),
        ("useSparseOptimization", C.c_int32),
        ("checkpointInterval", C.c_uint32),
        ("phase", C.c_int32),
        ("iter", C.c_uint32),
        ("nAtomsA", C.c_uint64),
        ("nAtomsP", C.c_uint64),
        ("statUpdates", C.c_uint32),
        ("fileBytes", C.c_uint64),
    ]
