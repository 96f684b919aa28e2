"""Loads libcogaps_b200.so (built in-tree by cogaps_b200/csrc/Makefile) and declares the C ABI.

There is no fallback: if the library is missing the import of anything that computes fails loudly,
and on a machine without an sm_100-class GPU every compute entry point returns CGB_ENODEVICE.
"""
import ctypes as C
import os

from ._abi import (CgbParams, CgbResult, CgbSamplerCounters, CgbReductionOrder, CgbRunOptions, CgbCheckpointInfo,
                   c_float_p, c_u32_p, c_u64_p, c_i32_p)

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libcogaps_b200.so")

# every symbol include/cogaps_b200.h declares
EXPORTS = [
    "cgb_last_error", "cgb_build_report", "cgb_set_device", "cgb_set_resident_share", "cgb_kernel_launch_count", "cgb_params_default",
    "cgb_run", "cgb_randstate_create", "cgb_randstate_set_tables", "cgb_randstate_get_tables",
    "cgb_randstate_next_seed", "cgb_randstate_destroy", "cgb_rng_create", "cgb_rng_uniform32",
    "cgb_rng_uniform32_range", "cgb_rng_uniform64_range", "cgb_rng_uniform", "cgb_rng_poisson",
    "cgb_rng_exponential", "cgb_rng_trunc_normal", "cgb_rng_trunc_gamma_upper", "cgb_rng_destroy",
    "cgb_sampler_create", "cgb_sampler_destroy", "cgb_sampler_set_uncertainty", "cgb_sampler_set_matrix",
    "cgb_sampler_set_annealing_temp", "cgb_sampler_sync", "cgb_sampler_extra_initialization",
    "cgb_sampler_update", "cgb_sampler_chisq", "cgb_sampler_n_atoms", "cgb_sampler_data_sparsity",
    "cgb_sampler_average_queue_length", "cgb_sampler_get_matrix", "cgb_sampler_shape", "cgb_sampler_lambda",
    "cgb_sampler_get_atoms", "cgb_sampler_get_ap_row", "cgb_sampler_alpha_parameters",
    "cgb_sampler_get_counters", "cgb_sampler_reset_counters", "cgb_sampler_set_kernel_timing", "cgb_sampler_set_persistent",
    "cgb_sampler_reduction_order", "cgb_reduction_order_for_length", "cgb_stats_create", "cgb_stats_destroy", "cgb_stats_update",
    "cgb_stats_update_a", "cgb_stats_update_p", "cgb_stats_update_pump", "cgb_stats_amean", "cgb_stats_asd",
    "cgb_stats_pmean", "cgb_stats_psd", "cgb_stats_pump_matrix", "cgb_stats_mean_pattern",
    "cgb_stats_mean_chisq", "cgb_sampler_device_matrix", "cgb_stats_device_sums", "cgb_run_set_tables",
    "cgb_debug_logf", "cgb_debug_host_logf", "cgb_debug_fastdiv", "cgb_run_file", "cgb_read_matrix_file",
    "cgb_run_ex", "cgb_sampler_serialize", "cgb_sampler_deserialize", "cgb_sampler_set_atoms", "cgb_stats_serialize",
    "cgb_stats_deserialize", "cgb_randstate_get_state", "cgb_randstate_set_state", "cgb_rng_get_state", "cgb_rng_set_state",
    "cgb_checkpoint_info_read", "cgb_checkpoint_rewrite", "cgb_read_matrix_csr",
    "cgb_debug_replay_generator", "cgb_debug_replay_message", "cgb_run_file_ex", "cgb_debug_running_sum", "cgb_write_matrix_csv", "cgb_result_write_files", "cgb_file_col_names", "cgb_debug_domain_fuzz", "cgb_release_device_cache",
    "cgb_sampler_set_update_mode", "cgb_sweep_reduction_order_for_length",
    "cgb_sampler_set_resident_share", "cgb_comm_get_unique_id", "cgb_comm_init", "cgb_comm_destroy", "cgb_allgather_rows", "cgb_allgather_device_rows",
]

_lib = None


class CogapsError(RuntimeError):
    def __init__(self, code, message):
        RuntimeError.__init__(self, "cogaps_b200 error %d: %s" % (code, message))
        self.code = code


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise OSError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                      "(or make -C cogaps_b200/csrc). There is no CPU fallback." % LIB_PATH)
    L = C.CDLL(LIB_PATH)
    vp = C.c_void_p
    L.cgb_last_error.restype = C.c_char_p
    L.cgb_build_report.restype = C.c_char_p
    L.cgb_kernel_launch_count.restype = C.c_uint64
    L.cgb_debug_host_logf.restype = C.c_float
    L.cgb_debug_host_logf.argtypes = [C.c_float]
    L.cgb_read_matrix_file.argtypes = [C.c_char_p, c_float_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
    L.cgb_run_file.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(CgbParams), C.POINTER(CgbResult)]
    L.cgb_debug_fastdiv.restype = C.c_uint64
    L.cgb_debug_fastdiv.argtypes = [C.c_uint64, C.c_uint64]
    L.cgb_run.argtypes = [c_float_p, C.c_uint32, C.c_uint32, C.c_int32, c_float_p, C.POINTER(CgbParams),
                          C.POINTER(CgbResult)]
    L.cgb_randstate_create.argtypes = [C.c_uint32, C.POINTER(vp)]
    L.cgb_randstate_set_tables.argtypes = [vp, c_float_p, c_float_p, c_float_p]
    L.cgb_randstate_get_tables.argtypes = [vp, c_float_p, c_float_p, c_float_p]
    L.cgb_randstate_next_seed.argtypes = [vp, c_u64_p]
    L.cgb_randstate_destroy.argtypes = [vp]
    L.cgb_randstate_destroy.restype = None
    L.cgb_rng_create.argtypes = [vp, C.POINTER(vp)]
    L.cgb_rng_destroy.argtypes = [vp]
    L.cgb_rng_destroy.restype = None
    L.cgb_rng_uniform32.argtypes = [vp, c_u32_p]
    L.cgb_rng_uniform32_range.argtypes = [vp, C.c_uint32, C.c_uint32, c_u32_p]
    L.cgb_rng_uniform64_range.argtypes = [vp, C.c_uint64, C.c_uint64, c_u64_p]
    L.cgb_rng_uniform.argtypes = [vp, c_float_p]
    L.cgb_rng_poisson.argtypes = [vp, C.c_double, c_i32_p]
    L.cgb_rng_exponential.argtypes = [vp, C.c_float, c_float_p]
    L.cgb_rng_trunc_normal.argtypes = [vp, C.c_float, C.c_float, C.c_float, C.c_float, c_float_p, c_i32_p]
    L.cgb_rng_trunc_gamma_upper.argtypes = [vp, C.c_float, C.c_float, c_float_p]
    L.cgb_sampler_create.argtypes = [c_float_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32, C.c_int32,
                                     C.c_float, C.c_float, C.POINTER(CgbParams), vp, C.POINTER(vp)]
    L.cgb_sampler_destroy.argtypes = [vp]
    L.cgb_sampler_destroy.restype = None
    L.cgb_sampler_set_uncertainty.argtypes = [vp, c_float_p, C.c_uint32, C.c_uint32, C.c_int32, C.c_int32,
                                              C.c_int32, C.POINTER(CgbParams)]
    L.cgb_sampler_set_matrix.argtypes = [vp, c_float_p]
    L.cgb_sampler_set_annealing_temp.argtypes = [vp, C.c_float]
    L.cgb_sampler_sync.argtypes = [vp, vp]
    L.cgb_sampler_extra_initialization.argtypes = [vp]
    L.cgb_sampler_update.argtypes = [vp, C.c_uint32, C.c_uint32]
    L.cgb_sampler_chisq.argtypes = [vp, c_float_p]
    L.cgb_sampler_n_atoms.argtypes = [vp, c_u64_p]
    L.cgb_sampler_data_sparsity.argtypes = [vp, c_float_p]
    L.cgb_sampler_average_queue_length.argtypes = [vp, c_float_p]
    L.cgb_sampler_get_matrix.argtypes = [vp, c_float_p]
    L.cgb_sampler_shape.argtypes = [vp, c_u32_p, c_u32_p, c_u32_p]
    L.cgb_sampler_lambda.argtypes = [vp, c_float_p, c_float_p]
    L.cgb_sampler_get_atoms.argtypes = [vp, c_u64_p, c_float_p, C.c_uint64, c_u64_p]
    L.cgb_sampler_get_ap_row.argtypes = [vp, C.c_uint32, c_float_p]
    L.cgb_sampler_alpha_parameters.argtypes = [vp, C.c_uint32, c_i32_p, c_u32_p, c_u32_p, c_u32_p, c_u32_p,
                                               c_float_p, c_float_p, c_float_p]
    L.cgb_sampler_get_counters.argtypes = [vp, C.POINTER(CgbSamplerCounters)]
    L.cgb_sampler_reset_counters.argtypes = [vp]
    L.cgb_sampler_set_kernel_timing.argtypes = [vp, C.c_int32]
    L.cgb_sampler_set_persistent.argtypes = [vp, C.c_int32]
    L.cgb_sampler_reduction_order.argtypes = [vp, C.POINTER(CgbReductionOrder)]
    L.cgb_reduction_order_for_length.argtypes = [C.c_uint32, C.POINTER(CgbReductionOrder)]
    L.cgb_sweep_reduction_order_for_length.argtypes = [C.c_uint32, C.POINTER(CgbReductionOrder)]
    L.cgb_sampler_set_update_mode.argtypes = [vp, C.c_int32]
    L.cgb_sampler_set_resident_share.argtypes = [vp, C.c_int32]
    L.cgb_comm_get_unique_id.argtypes = [vp]
    L.cgb_comm_init.argtypes = [vp, C.c_int32, C.c_int32, C.POINTER(vp)]
    L.cgb_comm_destroy.argtypes = [vp]
    L.cgb_comm_destroy.restype = None
    L.cgb_allgather_rows.argtypes = [vp, vp, c_u32_p, c_float_p, C.POINTER(C.c_double)]
    L.cgb_allgather_device_rows.argtypes = [vp, vp, C.c_uint64, C.c_uint32, c_u32_p, c_float_p, C.POINTER(C.c_double)]
    L.cgb_stats_create.argtypes = [C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(vp)]
    L.cgb_stats_destroy.argtypes = [vp]
    L.cgb_stats_destroy.restype = None
    for name in ("cgb_stats_update", "cgb_stats_update_a", "cgb_stats_update_p"):
        getattr(L, name).argtypes = [vp, vp, vp]
    L.cgb_stats_update_pump.argtypes = [vp, vp]
    for name in ("cgb_stats_amean", "cgb_stats_asd", "cgb_stats_pmean", "cgb_stats_psd",
                 "cgb_stats_pump_matrix", "cgb_stats_mean_pattern"):
        getattr(L, name).argtypes = [vp, c_float_p]
    L.cgb_stats_mean_chisq.argtypes = [vp, vp, c_float_p]
    L.cgb_sampler_device_matrix.argtypes = [vp, C.POINTER(vp), c_u64_p]
    L.cgb_stats_device_sums.argtypes = [vp, C.POINTER(vp), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp),
                                        c_u64_p, c_u64_p, c_u32_p]
    L.cgb_run_set_tables.argtypes = [c_float_p, c_float_p, c_float_p]
    L.cgb_debug_logf.argtypes = [c_float_p, c_float_p, C.c_uint32]
    L.cgb_run_ex.argtypes = [c_float_p, C.c_uint32, C.c_uint32, C.c_int32, c_float_p, C.POINTER(CgbParams),
                             C.POINTER(CgbRunOptions), C.POINTER(CgbResult)]
    L.cgb_sampler_serialize.argtypes = [vp, vp, C.c_uint64, c_u64_p]
    L.cgb_sampler_deserialize.argtypes = [vp, vp, C.c_uint64]
    L.cgb_sampler_set_atoms.argtypes = [vp, c_u64_p, c_float_p, C.c_uint64]
    L.cgb_stats_serialize.argtypes = [vp, vp, C.c_uint64, c_u64_p]
    L.cgb_stats_deserialize.argtypes = [vp, vp, C.c_uint64]
    L.cgb_randstate_get_state.argtypes = [vp, c_u64_p]
    L.cgb_randstate_set_state.argtypes = [vp, c_u64_p]
    L.cgb_rng_get_state.argtypes = [vp, c_u64_p]
    L.cgb_rng_set_state.argtypes = [vp, C.c_uint64]
    L.cgb_checkpoint_info_read.argtypes = [C.c_char_p, C.POINTER(CgbCheckpointInfo)]
    L.cgb_checkpoint_rewrite.argtypes = [C.c_char_p, C.c_char_p]
    L.cgb_read_matrix_csr.argtypes = [C.c_char_p, C.c_int32, c_u32_p, c_u32_p, c_u32_p, C.c_uint64, c_u32_p, c_float_p,
                                      C.c_uint64, c_u64_p]
    L.cgb_debug_replay_generator.argtypes = [c_float_p, C.c_uint32, C.c_uint32, C.POINTER(CgbParams), vp, C.c_uint64, c_u64_p]
    L.cgb_debug_replay_message.restype = C.c_char_p
    L.cgb_debug_domain_fuzz.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, c_u64_p]
    L.cgb_file_col_names.argtypes = [C.c_char_p, C.c_char_p, C.c_uint64, c_u64_p, c_u32_p]
    L.cgb_write_matrix_csv.argtypes = [C.c_char_p, c_float_p, C.c_uint32, C.c_uint32]
    L.cgb_result_write_files.argtypes = [C.c_char_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(CgbResult)]
    L.cgb_debug_running_sum.argtypes = [c_float_p, C.c_uint32, C.c_uint32, C.c_int32, c_float_p, c_u32_p]
    L.cgb_run_file_ex.argtypes = [C.c_char_p, C.c_char_p, C.POINTER(CgbParams), C.POINTER(CgbRunOptions), C.POINTER(CgbResult)]
    _lib = L
    return L


def check(rc):
    if rc != 0:
        raise CogapsError(rc, lib().cgb_last_error().decode(errors="replace"))
