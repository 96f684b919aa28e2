"""cogaps_b200 — the B200-native drop-in for CoGAPS's Gibbs-sampler hot path.

Layout: csrc/ holds the CUDA kernels and the C ABI (libcogaps_b200.so, declared in include/cogaps_b200.h);
this package is the thin host-side mirror of the reference's R/C++ interface for that path.
"""
from .api import (CoGAPS, CogapsParams, CogapsResult, gaps_run, gaps_run_file, read_matrix_file,  # noqa: F401
                  checkpoint_info, checkpoint_rewrite, read_matrix_csr, write_matrix_csv, write_result_files,
                  buildReport, checkpointsEnabled, compiledWithOpenMPSupport, getFileInfo)
from .sampler import GapsRandomState, GapsRng, GibbsSampler, GapsStatistics, Comm  # noqa: F401
from ._lib import CogapsError, LIB_PATH, lib  # noqa: F401
