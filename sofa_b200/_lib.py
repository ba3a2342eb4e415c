"""ctypes binding of libsofa_b200.so (include/sofa_b200.h).  Fails loudly when the CUDA
library is missing or no GPU is present: there is no CPU fallback."""
import ctypes
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libsofa_b200.so")

SOFA_OK, SOFA_ERR_CUDA, SOFA_ERR_ARG, SOFA_ERR_ROW, SOFA_ERR_STATE, SOFA_ERR_CAP = 0, -1, -2, -3, -4, -5

SOFA_DEBUG_NAME_HASH_MASK = 1

EXPORTS = [
    "sofa_b200_abi_version", "sofa_b200_last_error", "sofa_b200_error_row", "sofa_b200_create",
    "sofa_b200_destroy", "sofa_b200_host_alloc", "sofa_b200_host_free", "sofa_b200_upload",
    "sofa_b200_parse", "sofa_b200_row_flags", "sofa_b200_emit_csv", "sofa_b200_set_filters",
    "sofa_b200_viz_mangled", "sofa_b200_set_demangled", "sofa_b200_emit_json", "sofa_b200_emit_json_series", "sofa_b200_set_debug", "sofa_b200_iter_profile", "sofa_b200_synth_range", "sofa_b200_shard_export", "sofa_b200_shard_reserve", "sofa_b200_shard_commit", "sofa_b200_set_global_flags", "sofa_b200_net_triples", "sofa_b200_set_net_filters", "sofa_b200_mpstat_profile",
    "sofa_b200_side_quantiles", "sofa_b200_side_int_sums", "sofa_b200_nvsmi_roi", "sofa_b200_fetch_range", "sofa_b200_emit_block_end", "sofa_b200_fetch",
    "sofa_b200_get_column", "sofa_b200_gpu_profile", "sofa_b200_top_k", "sofa_b200_cpu_profile",
    "sofa_b200_set_profiling", "sofa_b200_get_stats", "sofa_b200_reset_stats", "sofa_b200_kernel_name",
    "sofa_b200_synth", "sofa_b200_text_fetch", "sofa_b200_mark", "sofa_b200_elapsed_ms",
    "sofa_b200_upload_async", "sofa_b200_fetch_async", "sofa_b200_wait_fetch", "sofa_b200_sync", "sofa_b200_set_output_arena", "sofa_b200_emit_comm_csv", "sofa_b200_window_reduce",
    "sofa_b200_side2_preprocess", "sofa_b200_side2_dev_table", "sofa_b200_side2_set_devs", "sofa_b200_side2_emit",
    "sofa_b200_diskstat_profile", "sofa_b200_blktrace_latency",
    "sofa_b200_upload_chunk", "sofa_b200_hold_output", "sofa_b200_hold_fetch", "sofa_b200_hold_release",
    "sofa_b200_side_preprocess", "sofa_b200_side_emit_csv", "sofa_b200_side_emit_json", "sofa_b200_side_emit_aux", "sofa_b200_nvsmi_profile", "sofa_b200_net_profile",
]


class IngestParams(ctypes.Structure):
    _fields_ = [("time_base", ctypes.c_double), ("t_offset", ctypes.c_double),
                ("absolute_timestamp", ctypes.c_int32), ("perf_has_cycles", ctypes.c_int32),
                ("n_mhz", ctypes.c_int32), ("n_cudaproc", ctypes.c_int32),
                ("mhz_xp", ctypes.c_void_p), ("mhz_fp", ctypes.c_void_p),
                ("ts_rescale", ctypes.c_double), ("dt_rescale", ctypes.c_double),
                ("payload_unit", ctypes.c_double), ("idx", ctypes.c_int32 * 9),
                ("api_mode", ctypes.c_int32)]


class GpuProfile(ctypes.Structure):
    _fields_ = [("gpu_time", ctypes.c_double), ("kernel_time", ctypes.c_double),
                ("nccl_time", ctypes.c_double), ("num_gpus", ctypes.c_int64),
                ("n_gpus_max", ctypes.c_int64), ("payload", ctypes.c_double * 4),
                ("bandwidth", ctypes.c_double * 4), ("duration", ctypes.c_double * 4),
                ("accum", ctypes.c_double * 256), ("accum_count", ctypes.c_double * 256),
                ("accum_bw", ctypes.c_double * 256), ("accum_bw_count", ctypes.c_double * 256),
                ("per_gpu_id", ctypes.c_int64 * 64), ("per_gpu_time", ctypes.c_double * 64)]


class Window(ctypes.Structure):
    _fields_ = [("mp_count", ctypes.c_uint32), ("usr_max", ctypes.c_int32), ("sys_max", ctypes.c_int32),
                ("idl_max", ctypes.c_int32), ("iow_max", ctypes.c_int32), ("cpu_max", ctypes.c_int32),
                ("cpu_min", ctypes.c_int32), ("usr_sum", ctypes.c_int64), ("sys_sum", ctypes.c_int64),
                ("irq_sum", ctypes.c_int64), ("nv_count", ctypes.c_uint32), ("nv_ndev", ctypes.c_uint32),
                ("nv_max", ctypes.c_double), ("nv_min", ctypes.c_double), ("nv_sum", ctypes.c_int64),
                ("tx_count", ctypes.c_uint32), ("rx_count", ctypes.c_uint32), ("tx_sum", ctypes.c_double),
                ("rx_sum", ctypes.c_double)]


class NetPair(ctypes.Structure):
    _fields_ = [("src", ctypes.c_int64), ("dst", ctypes.c_int64), ("payload_sum", ctypes.c_int64), ("count", ctypes.c_int64)]


class NetTriple(ctypes.Structure):
    _fields_ = [("src", ctypes.c_int64), ("dst", ctypes.c_int64), ("payload", ctypes.c_int64), ("count", ctypes.c_int64)]


class ShardDesc(ctypes.Structure):
    _fields_ = [("n_rows", ctypes.c_uint64), ("text_bytes", ctypes.c_uint64), ("data_off", ctypes.c_uint64),
                ("text", ctypes.c_void_p), ("cols", ctypes.c_void_p * 16), ("col_elem", ctypes.c_uint32 * 16),
                ("n_cols", ctypes.c_int32), ("pad", ctypes.c_int32)]


class IterStats(ctypes.Structure):
    _fields_ = [("n_rows", ctypes.c_uint64), ("first_ts", ctypes.c_double), ("last_ts", ctypes.c_double),
                ("gpu_time", ctypes.c_double), ("fw_time", ctypes.c_double), ("bw_time", ctypes.c_double),
                ("gemm_time", ctypes.c_double), ("copy_name_time", ctypes.c_double), ("allreduce_time", ctypes.c_double),
                ("payload", ctypes.c_int64), ("streams", ctypes.c_uint32), ("pad", ctypes.c_uint32)]


class NvsmiGroup(ctypes.Structure):
    _fields_ = [("dev", ctypes.c_int32), ("event", ctypes.c_int32), ("count", ctypes.c_uint64), ("sum", ctypes.c_int64),
                ("q_lo", ctypes.c_double * 3), ("q_hi", ctypes.c_double * 3)]


class Side2Params(ctypes.Structure):
    _fields_ = [("time_base", ctypes.c_double), ("strace_min_time", ctypes.c_double), ("pid", ctypes.c_int64),
                ("category", ctypes.c_int64), ("category_text", ctypes.c_char_p), ("absolute_timestamp", ctypes.c_int32),
                ("plot_ratio", ctypes.c_int32), ("category_is_str", ctypes.c_int32), ("category_len", ctypes.c_int32)]


class Side2Info(ctypes.Structure):
    _fields_ = [("n_lines", ctypes.c_uint64), ("n_entries", ctypes.c_uint64), ("n_kept", ctypes.c_uint64),
                ("n_devs", ctypes.c_uint64), ("total", ctypes.c_double), ("last_span_a", ctypes.c_uint32),
                ("last_span_b", ctypes.c_uint32), ("flags", ctypes.c_int32), ("pad", ctypes.c_int32)]


SIDE2_DEV_CAP = 1024


class DiskGroup(ctypes.Structure):
    _fields_ = [("slot", ctypes.c_uint32), ("pad", ctypes.c_uint32), ("count", ctypes.c_uint64),
                ("q_lo", ctypes.c_double * 9), ("q_hi", ctypes.c_double * 9), ("mean", ctypes.c_double * 7)]


class SideParams(ctypes.Structure):
    _fields_ = [("time_base", ctypes.c_double), ("absolute_timestamp", ctypes.c_int32), ("reserved", ctypes.c_int32),
                ("local_utc_offset", ctypes.c_int64), ("nvsmi_tz_add", ctypes.c_double)]


class Stats(ctypes.Structure):
    _fields_ = [("launches", ctypes.c_uint64 * 32), ("ms", ctypes.c_double * 32),
                ("h2d_bytes", ctypes.c_uint64), ("d2h_bytes", ctypes.c_uint64)]


class SofaB200Error(RuntimeError):
    def __init__(self, code, msg, row=None, row_code=None):
        RuntimeError.__init__(self, "libsofa_b200 error %d: %s" % (code, msg))
        self.code, self.row, self.row_code = code, row, row_code


_lib = None


def load_lib():
    """Load libsofa_b200.so; raise if it has not been built (python __graft_entry__.py)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise ImportError("%s is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                          "sofa_b200 has no CPU fallback" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    u64p = ctypes.POINTER(ctypes.c_uint64)
    u32p = ctypes.POINTER(ctypes.c_uint32)
    vp = ctypes.c_void_p
    lib.sofa_b200_last_error.restype = ctypes.c_char_p
    lib.sofa_b200_kernel_name.restype = ctypes.c_char_p
    lib.sofa_b200_host_alloc.restype = ctypes.c_void_p
    lib.sofa_b200_host_alloc.argtypes = [ctypes.c_uint64]
    lib.sofa_b200_host_free.argtypes = [vp]
    lib.sofa_b200_create.argtypes = [ctypes.POINTER(vp), ctypes.c_int]
    lib.sofa_b200_destroy.argtypes = [vp]
    lib.sofa_b200_error_row.argtypes = [vp, u64p, ctypes.POINTER(ctypes.c_int32)]
    lib.sofa_b200_upload.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64, ctypes.c_uint64]
    lib.sofa_b200_parse.argtypes = [vp, ctypes.c_int, ctypes.POINTER(IngestParams), u64p]
    lib.sofa_b200_row_flags.argtypes = [vp, ctypes.c_int, ctypes.c_uint32, u32p]
    lib.sofa_b200_emit_csv.argtypes = [vp, ctypes.c_int, ctypes.c_uint32, u64p]
    lib.sofa_b200_emit_comm_csv.argtypes = [vp, u64p]
    lib.sofa_b200_set_filters.argtypes = [vp, ctypes.c_int, ctypes.c_uint32, ctypes.c_int32,
                                          ctypes.POINTER(ctypes.c_char_p), u64p]
    lib.sofa_b200_viz_mangled.argtypes = [vp, ctypes.c_uint32, u64p, u32p]
    lib.sofa_b200_set_demangled.argtypes = [vp, ctypes.c_uint32, ctypes.c_char_p, u64p, ctypes.c_char_p, u64p]
    lib.sofa_b200_emit_json.argtypes = [vp, ctypes.c_int, ctypes.c_uint32, ctypes.c_int32, u64p, u64p]
    lib.sofa_b200_emit_json_series.argtypes = [vp, ctypes.c_int, ctypes.c_uint32, ctypes.c_int32, ctypes.POINTER(ctypes.c_char_p), u64p, u64p, u64p]
    lib.sofa_b200_fetch.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64]
    lib.sofa_b200_set_debug.argtypes = [vp, ctypes.c_int, ctypes.c_uint64]
    lib.sofa_b200_iter_profile.argtypes = [vp, ctypes.c_int32, ctypes.POINTER(ctypes.c_double), ctypes.c_uint32, ctypes.POINTER(IterStats)]
    lib.sofa_b200_shard_export.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ShardDesc)]
    lib.sofa_b200_shard_reserve.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.POINTER(ShardDesc)]
    lib.sofa_b200_shard_commit.argtypes = [vp, ctypes.c_int, ctypes.c_int32, u64p, u64p]
    lib.sofa_b200_set_global_flags.argtypes = [vp, ctypes.c_int, ctypes.c_uint32, ctypes.c_uint32]
    dp = ctypes.POINTER(ctypes.c_double)
    i64p = ctypes.POINTER(ctypes.c_int64)
    lib.sofa_b200_net_triples.argtypes = [vp, ctypes.c_uint32, ctypes.POINTER(NetTriple), ctypes.c_uint32, u32p]
    lib.sofa_b200_set_net_filters.argtypes = [vp, ctypes.c_uint32, ctypes.c_int64, ctypes.c_int32, i64p, u64p]
    lib.sofa_b200_mpstat_profile.argtypes = [vp, dp, ctypes.c_uint32, u32p]
    lib.sofa_b200_side_quantiles.argtypes = [vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_double, dp, u64p, dp, dp]
    lib.sofa_b200_side_int_sums.argtypes = [vp, ctypes.c_int, i64p, u64p]
    lib.sofa_b200_nvsmi_roi.argtypes = [vp, dp, dp]
    lib.sofa_b200_fetch_range.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, vp]
    lib.sofa_b200_emit_block_end.argtypes = [vp, ctypes.c_uint64, u64p]
    lib.sofa_b200_get_column.argtypes = [vp, ctypes.c_int, ctypes.c_int, vp, ctypes.c_uint64]
    lib.sofa_b200_gpu_profile.argtypes = [vp, ctypes.POINTER(GpuProfile), ctypes.c_int]
    lib.sofa_b200_top_k.argtypes = [vp, ctypes.c_int, ctypes.c_uint32, ctypes.POINTER(ctypes.c_double), u64p, u32p]
    lib.sofa_b200_cpu_profile.argtypes = [vp, ctypes.c_uint32, ctypes.POINTER(ctypes.c_double), u64p,
                                          ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), u64p, u32p]
    lib.sofa_b200_upload_async.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64, ctypes.c_uint64]
    lib.sofa_b200_fetch_async.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64, ctypes.POINTER(ctypes.c_int32)]
    lib.sofa_b200_wait_fetch.argtypes = [vp, ctypes.c_int32]
    lib.sofa_b200_sync.argtypes = [vp]
    lib.sofa_b200_set_output_arena.argtypes = [vp, ctypes.c_uint64]
    lib.sofa_b200_mark.argtypes = [vp, ctypes.c_int]
    lib.sofa_b200_elapsed_ms.argtypes = [vp, ctypes.POINTER(ctypes.c_double)]
    lib.sofa_b200_window_reduce.argtypes = [vp, ctypes.POINTER(ctypes.c_double), ctypes.c_uint32, ctypes.c_double,
                                            ctypes.POINTER(Window), u32p]
    lib.sofa_b200_net_profile.argtypes = [vp, ctypes.c_uint32, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int64),
                                          ctypes.POINTER(NetPair), ctypes.c_uint32, u32p]
    lib.sofa_b200_nvsmi_profile.argtypes = [vp, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.POINTER(NvsmiGroup),
                                            ctypes.c_uint32, u32p, ctypes.POINTER(ctypes.c_int32), u32p,
                                            ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double), u64p]
    lib.sofa_b200_side_preprocess.argtypes = [vp, ctypes.c_int, ctypes.POINTER(SideParams), u64p, u64p,
                                              ctypes.POINTER(ctypes.c_int32)]
    lib.sofa_b200_side_emit_csv.argtypes = [vp, ctypes.c_int, ctypes.c_int, u64p, u64p]
    lib.sofa_b200_side_emit_json.argtypes = [vp, ctypes.c_int, ctypes.c_int, u64p, u64p]
    lib.sofa_b200_side_emit_aux.argtypes = [vp, ctypes.c_int, u64p, u64p]
    lib.sofa_b200_side2_preprocess.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64, ctypes.POINTER(Side2Params), ctypes.POINTER(Side2Info)]
    lib.sofa_b200_side2_dev_table.argtypes = [vp, vp, ctypes.c_uint32]
    lib.sofa_b200_side2_set_devs.argtypes = [vp, vp, vp, ctypes.c_uint32]
    lib.sofa_b200_side2_emit.argtypes = [vp, ctypes.c_int, ctypes.c_int, u64p, u64p]
    lib.sofa_b200_upload_chunk.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, vp, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64]
    lib.sofa_b200_hold_output.argtypes = [vp, ctypes.c_int, ctypes.POINTER(ctypes.c_int32)]
    lib.sofa_b200_hold_fetch.argtypes = [vp, ctypes.c_int32, ctypes.c_uint64, ctypes.c_uint64, vp]
    lib.sofa_b200_hold_release.argtypes = [vp, ctypes.c_int32]
    lib.sofa_b200_diskstat_profile.argtypes = [vp, vp, ctypes.c_uint64, ctypes.POINTER(DiskGroup), ctypes.c_uint32, u32p, vp]
    lib.sofa_b200_blktrace_latency.argtypes = [vp, vp, ctypes.c_uint64, u64p, ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_double)]
    lib.sofa_b200_set_profiling.argtypes = [vp, ctypes.c_int]
    lib.sofa_b200_get_stats.argtypes = [vp, ctypes.POINTER(Stats)]
    lib.sofa_b200_reset_stats.argtypes = [vp]
    lib.sofa_b200_synth.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, u64p]
    lib.sofa_b200_synth_range.argtypes = [vp, ctypes.c_int, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_uint64, ctypes.c_int, ctypes.c_int, u64p]
    lib.sofa_b200_text_fetch.argtypes = [vp, ctypes.c_int, vp, ctypes.c_uint64]
    if lib.sofa_b200_abi_version() != 2:
        raise ImportError("libsofa_b200.so ABI version mismatch")
    _lib = lib
    return lib
