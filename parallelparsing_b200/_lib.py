"""ctypes binding of libppb200.so (include/ppb200.h).  No fallback: if the CUDA
library is missing or there is no sm_100 device the calls raise."""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libppb200.so")

WINSIZE = 32768

PP_OK = 0
PP_E_CUDA, PP_E_NO_DEVICE, PP_E_ARG, PP_E_IO, PP_E_RECORD_TOO_LONG, PP_E_FORMAT, PP_E_UNSUPPORTED = -100, -101, -102, -103, -104, -105, -106
PP_INDEX_LIFT_RECORD_CAP = 1
PP_JOB_STRICT, PP_JOB_ZEROCOPY, PP_JOB_PIPELINE, PP_JOB_COMPACT_WINDOWS = 1, 2, 4, 8


class PPPoint(C.Structure):
    _fields_ = [("output", C.c_int64), ("input", C.c_int64), ("bits", C.c_int32), ("offset_len", C.c_int32),
                ("window", C.POINTER(C.c_uint8)), ("offset", C.POINTER(C.c_uint8))]


class PPJobInfo(C.Structure):
    _fields_ = [("first_chunk", C.c_int32), ("n_chunks", C.c_int32), ("total_records", C.c_int64),
                ("total_bytes", C.c_int64), ("scanned_bytes", C.c_int64), ("compressed_bytes", C.c_int64),
                ("h2d_bytes", C.c_int64), ("d2h_bytes", C.c_int64), ("status", C.c_int32),
                ("exact_chunks", C.c_int32), ("upload_ms", C.c_float), ("inflate_ms", C.c_float),
                ("scan_ms", C.c_float), ("parse_ms", C.c_float), ("download_ms", C.c_float),
                ("launches", C.c_int32)]


class PPChunkInfo(C.Structure):
    _fields_ = [("status", C.c_int32), ("prefix_len", C.c_int32), ("inflated", C.c_int64), ("records", C.c_int64),
                ("record_base", C.c_int64), ("parse_end", C.c_uint32), ("flags", C.c_uint32)]


class PPMultiInfo(C.Structure):
    _fields_ = [("n_parts", C.c_int32), ("n_chunks", C.c_int32), ("total_records", C.c_int64),
                ("total_bytes", C.c_int64), ("compressed_bytes", C.c_int64), ("status", C.c_int32), ("pad", C.c_int32)]


class PPPairInfo(C.Structure):
    _fields_ = [("n_parts", C.c_int32), ("topup_chunks", C.c_int32), ("records_r1", C.c_int64), ("records_r2", C.c_int64),
                ("pairs", C.c_int64), ("status", C.c_int32), ("pad", C.c_int32)]


class ZException(Exception):
    """Interop/Conventions.cs:33-41 — carries the ZResult-compatible code."""

    def __init__(self, code, what=""):
        self.Code = code
        msg = lib().pp_strerror(code).decode() if _lib is not None else str(code)
        super().__init__(f"{what}: {msg} ({code})" if what else f"{msg} ({code})")


_lib = None

# every symbol include/ppb200.h declares: (name, restype, argtypes)
_p, _i32, _i64, _u32, _sz = C.c_void_p, C.c_int32, C.c_int64, C.c_uint32, C.c_size_t
_PP = C.POINTER(_p)
class PPCreateStats(C.Structure):
    """include/ppb200.h pp_create_stats."""
    _fields_ = [(n, C.c_float) for n in ("h2d_ms", "scan_ms", "scan_kernel_ms", "plan_ms", "inflate_ms", "chain_ms",
                                         "resolve_ms", "count_crc_ms", "gather_ms", "total_ms")] + \
               [("blocks", C.c_int64), ("total_out", C.c_int64), ("segments", C.c_int32), ("scan_passes", C.c_int32),
                ("points", C.c_int32), ("pad", C.c_int32)]


SYMBOLS = [
    ("pp_abi_version", C.c_int, []),
    ("pp_strerror", C.c_char_p, [C.c_int]),
    ("pp_index_create", C.c_int, [_p, _sz, _u32, _u32, _PP]),
    ("pp_index_create_file", C.c_int, [C.c_char_p, _u32, _u32, _PP]),
    ("pp_index_new", C.c_int, [_PP]),
    ("pp_index_add_point", C.c_int, [_p, _i32, _i64, _i64, _u32, _p, _p, _i32]),
    ("pp_index_add", C.c_int, [_p, _i32, _i64, _i64, _p, _p, _i32]),
    ("pp_index_serialize", C.c_int, [_p, C.c_char_p]),
    ("pp_index_serialize_v1", C.c_int, [_p, C.c_char_p]),
    ("pp_index_deserialize", C.c_int, [C.c_char_p, _PP]),
    ("pp_index_count", _i32, [_p]),
    ("pp_index_chunk_max_bytes", _i32, [_p]),
    ("pp_index_point", C.c_int, [_p, _i32, C.POINTER(PPPoint)]),
    ("pp_index_free", None, [_p]),
    ("pp_open", C.c_int, [_i32, _PP]),
    ("pp_close", None, [_p]),
    ("pp_host_alloc", C.c_int, [_sz, _PP]),
    ("pp_host_free", None, [_p]),
    ("pp_host_register", C.c_int, [_p, _sz]),
    ("pp_host_unregister", None, [_p]),
    ("pp_extract", _i64, [_p, _p, _i64, _p, _i32, _p, _i64]),
    ("pp_parse", _i64, [_p, _p, _i64, _p, _i64, _p, _i64, C.POINTER(_u32)]),
    ("pp_job_create", C.c_int, [_p, _p, _sz, _i32, _i32, _u32, _PP]),
    ("pp_job_upload", C.c_int, [_p, _p]),
    ("pp_job_file_range", C.c_int, [_p, C.POINTER(_i64), C.POINTER(_i64)]),
    ("pp_job_upload_range", C.c_int, [_p, _p, _i64, _i64]),
    ("pp_job_execute", C.c_int, [_p]),
    ("pp_job_execute_to_host", C.c_int, [_p, _p, _i64]),
    ("pp_job_download", C.c_int, [_p]),
    ("pp_job_info_get", C.c_int, [_p, C.POINTER(PPJobInfo)]),
    ("pp_job_chunk_info", C.c_int, [_p, _i32, C.POINTER(PPChunkInfo)]),
    ("pp_job_fetch_line_starts", C.c_int, [_p, _p, _p, _p, _p]),
    ("pp_job_fetch_chunk", C.c_int, [_p, _i32, _p, _i64]),
    ("pp_job_fetch_bytes", C.c_int, [_p, _p, _i64]),
    ("pp_job_device_ptrs", C.c_int, [_p, _PP, _PP, _PP, _PP, _PP, _PP]),
    ("pp_job_base_histogram", C.c_int, [_p, _p]),
    ("pp_job_count_pattern", C.c_int, [_p, _p, _i32, _p]),
    ("pp_job_digests", C.c_int, [_p, _p, _p]),
    ("pp_job_free", None, [_p]),
    ("pp_decompress_all", C.c_int, [_p, _p, _p, _sz, _i32, _i32, _u32, _PP]),
    ("pp_scan_blocks", C.c_int, [_p, _p, _sz, _i64, _p, _p, _i64, C.POINTER(_i64), C.POINTER(_i64), C.POINTER(_i64),
                                 C.POINTER(C.c_float), C.POINTER(_i32)]),
    ("pp_index_create_gpu", C.c_int, [_p, _p, _sz, _u32, _u32, _PP, C.POINTER(PPCreateStats)]),
    ("pp_partition_chunks", C.c_int, [_p, _i32, _p, _p]),
    ("pp_decompress_all_multi", C.c_int, [_p, _i32, _p, _p, _sz, _u32, _PP]),
    ("pp_multi_info_get", C.c_int, [_p, C.POINTER(PPMultiInfo)]),
    ("pp_multi_part", C.c_int, [_p, _i32, _PP, C.POINTER(_i32), C.POINTER(_i64)]),
    ("pp_multi_free", None, [_p]),
    ("pp_release_cached_contexts", None, []),
    ("pp_pair_decompress_all", C.c_int, [_p, _i32, _p, _p, _sz, _p, _p, _sz, _u32, _PP]),
    ("pp_pair_info_get", C.c_int, [_p, C.POINTER(PPPairInfo)]),
    ("pp_pair_part", C.c_int, [_p, _i32, _PP, C.POINTER(_i64), C.POINTER(_i32)]),
    ("pp_pair_part_r2", C.c_int, [_p, _i32, _i32, _PP, C.POINTER(_i64)]),
    ("pp_pair_locate", C.c_int, [_p, _i32, _i64, C.POINTER(_i32), C.POINTER(_i64)]),
    ("pp_pair_free", None, [_p]),
]


def lib():
    """Load libppb200.so; raise (never fall back) when it has not been built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise RuntimeError(
                f"{LIB_PATH} is missing: build it with `python -m parallelparsing_b200.build` "
                "(nvcc, sm_100a). There is no CPU fallback.")
        L = C.CDLL(LIB_PATH)
        for name, rt, at in SYMBOLS:
            f = getattr(L, name)
            f.restype = rt
            f.argtypes = at
        if L.pp_abi_version() != 1:
            raise RuntimeError("libppb200.so ABI version mismatch")
        _lib = L
    return _lib


def check(rc, what=""):
    if rc < 0:
        raise ZException(rc, what)
    return rc
