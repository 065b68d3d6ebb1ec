"""ctypes binding of libfastq_b200.so -- the C ABI declared in include/fastq_b200.h.

The CUDA library is the only implementation of the hot path: loading fails loudly when the
shared object is missing (no CPU fallback, no oracle import).
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.environ.get("FQB_LIB") or os.path.join(_HERE, "libfastq_b200.so")   # FQB_LIB: A/B builds (tools/)
ABI_VERSION = int(os.environ.get("FQB_ABI", "3"))   # (FQB_ABI: A/B runs against an older build, tools/)

# status codes (include/fastq_b200.h)
OK, E_HEADER, E_SEP, E_LENGTH, E_TOO_LONG, E_TRUNCATED, E_IO, E_PHASE, E_RETRY = range(9)
E_ARG, E_STATE, E_NOMEM, E_CANCELLED, E_CUDA, E_NCCL = 50, 51, 52, 53, 100, 101
MAX_WORLD, COMM_ID_BYTES = 64, 128
KEEP_ALL, KEEP_DNA, KEEP_DNAN = 0, 1, 2
F_HIST, F_INDEX, F_LINE_START, F_EOF, F_FRONT16, F_INFER_START, F_PARTIAL = 0x01, 0x02, 0x04, 0x08, 0x10, 0x20, 0x40
MAX_RECORD_BYTES = 68 * 1024
SYNTH_SEED = 0xFA57A11CE5EED001
NO_OFFSET = 0xFFFFFFFFFFFFFFFF

# every symbol include/fastq_b200.h declares (tests check that the .so exports all of them)
SYMBOLS = [
    "fqb_abi_version", "fqb_stats_words", "fqb_stats_len_hist_off", "fqb_stats_base_hist_off",
    "fqb_stats_qual_hist_off", "fqb_create", "fqb_destroy", "fqb_strerror", "fqb_last_error",
    "fqb_parse_device", "fqb_count_lines_device", "fqb_fetch_line_count", "fqb_fetch",
    "fqb_device_stats", "fqb_device_result", "fqb_launch_count", "fqb_last_scan_ms", "fqb_last_index_ms", "fqb_parse_host",
    "fqb_stream_begin", "fqb_stream_acquire", "fqb_stream_submit", "fqb_stream_finish",
    "fqb_host_alloc", "fqb_host_free", "fqb_synth_fixed_device", "fqb_synth_var_device",
    "fqb_synth_var_sizes_device", "fqb_filter_device", "fqb_fetch_filter", "fqb_last_path",
    "fqb_comm_unique_id", "fqb_comm_init", "fqb_comm_destroy", "fqb_comm_rank", "fqb_comm_world", "fqb_allreduce",
    "fqb_fetch_reduced", "fqb_device_exchange", "fqb_exchange_words",
    "fqb_batch_begin", "fqb_batch_close", "fqb_next_batch", "fqb_release_batch", "fqb_batch_cancel", "fqb_batch_end",
]


class Config(C.Structure):
    _fields_ = [("abi_version", C.c_uint32), ("device", C.c_int32), ("max_len", C.c_uint32),
                ("reserved0", C.c_uint32), ("slot_bytes", C.c_uint64), ("n_slots", C.c_uint32),
                ("reserved1", C.c_uint32)]


class Shard(C.Structure):
    _fields_ = [("d_bytes", C.c_void_p), ("n_own", C.c_uint64), ("n_avail", C.c_uint64),
                ("stream_offset", C.c_uint64), ("line_base", C.c_uint64), ("flags", C.c_uint32),
                ("reserved", C.c_uint32), ("d_index", C.c_void_p), ("index_cap", C.c_uint64)]


class Result(C.Structure):
    _fields_ = [("status", C.c_int32), ("finished", C.c_int32), ("n_records", C.c_uint64),
                ("n_lines", C.c_uint64), ("err_offset", C.c_uint64), ("tail_offset", C.c_uint64),
                ("line_phase", C.c_uint32), ("reserved", C.c_uint32)]


class Batch(C.Structure):
    _fields_ = [("bytes", C.c_void_p), ("n_bytes", C.c_uint64), ("n_avail", C.c_uint64), ("stream_offset", C.c_uint64),
                ("line_ends", C.c_void_p), ("n_records", C.c_uint64), ("first_record", C.c_uint64),
                ("err_offset", C.c_uint64), ("token", C.c_uint64), ("status", C.c_int32), ("last", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile the library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    src_dir = os.path.join(_HERE, "csrc")
    srcs = [os.path.join(src_dir, f) for f in ("fq_scan.cu", "fq_stream.cu", "fq_kernels.cu", "fq_filter.cu", "fq_api.cu", "fq_common.cuh", "fq_device.cuh", "fq_hist.cuh")]
    srcs.append(os.path.join(os.path.dirname(_HERE), "include", "fastq_b200.h"))
    stale = (not os.path.exists(SO_PATH)) or any(
        os.path.getmtime(p) > os.path.getmtime(SO_PATH) for p in srcs if os.path.exists(p))
    if force or stale:
        subprocess.check_call(["make", "-C", src_dir] + (["-B"] if force else []),
                              stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    return SO_PATH


_lib = None


def lib():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise RuntimeError(
            f"{SO_PATH} is missing: the CUDA extension is the only implementation of the FASTQ "
            "hot path (no CPU fallback). Build it with `python -c 'import __graft_entry__ as g; "
            "g.build()'` or `make -C fastq_rs_b200/csrc`.")
    L = C.CDLL(SO_PATH)
    vp, u64, u32, i32 = C.c_void_p, C.c_uint64, C.c_uint32, C.c_int
    L.fqb_abi_version.restype = u32
    for n in ("fqb_stats_words", "fqb_stats_len_hist_off", "fqb_stats_base_hist_off",
              "fqb_stats_qual_hist_off"):
        getattr(L, n).argtypes = [u32]
        getattr(L, n).restype = C.c_size_t
    L.fqb_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    L.fqb_create.restype = i32
    L.fqb_destroy.argtypes = [vp]
    L.fqb_destroy.restype = None
    L.fqb_strerror.argtypes = [i32]
    L.fqb_strerror.restype = C.c_char_p
    L.fqb_last_error.argtypes = [vp]
    L.fqb_last_error.restype = C.c_char_p
    L.fqb_parse_device.argtypes = [vp, C.POINTER(Shard), vp]
    L.fqb_parse_device.restype = i32
    L.fqb_count_lines_device.argtypes = [vp, vp, u64, vp]
    L.fqb_count_lines_device.restype = i32
    L.fqb_fetch_line_count.argtypes = [vp, vp, C.POINTER(u64)]
    L.fqb_fetch_line_count.restype = i32
    L.fqb_fetch.argtypes = [vp, vp, C.POINTER(Result), vp]
    L.fqb_fetch.restype = i32
    L.fqb_device_stats.argtypes = [vp]
    L.fqb_device_stats.restype = vp
    L.fqb_device_result.argtypes = [vp]
    L.fqb_device_result.restype = vp
    L.fqb_launch_count.argtypes = [vp]
    L.fqb_launch_count.restype = u64
    L.fqb_last_index_ms.argtypes = [vp]
    L.fqb_last_index_ms.restype = C.c_float
    L.fqb_last_scan_ms.argtypes = [vp]
    L.fqb_last_scan_ms.restype = C.c_float
    L.fqb_parse_host.argtypes = [vp, vp, u64, u64, u32, C.POINTER(Result), vp, vp, u64, C.POINTER(u64)]
    L.fqb_parse_host.restype = i32
    L.fqb_stream_begin.argtypes = [vp, u32]
    L.fqb_stream_begin.restype = i32
    L.fqb_stream_acquire.argtypes = [vp, C.POINTER(vp), C.POINTER(u64)]
    L.fqb_stream_acquire.restype = i32
    L.fqb_stream_submit.argtypes = [vp, u64]
    L.fqb_stream_submit.restype = i32
    L.fqb_stream_finish.argtypes = [vp, C.POINTER(Result), vp]
    L.fqb_stream_finish.restype = i32
    L.fqb_host_alloc.argtypes = [u64, C.POINTER(vp)]
    L.fqb_host_alloc.restype = i32
    L.fqb_host_free.argtypes = [vp]
    L.fqb_host_free.restype = None
    L.fqb_synth_fixed_device.argtypes = [vp, u64, u64, u32, u64, vp]
    L.fqb_synth_fixed_device.restype = i32
    L.fqb_synth_var_device.argtypes = [vp, vp, u64, u64, u64, vp]
    L.fqb_synth_var_device.restype = i32
    L.fqb_synth_var_sizes_device.argtypes = [vp, u64, u64, u64, vp]
    L.fqb_synth_var_sizes_device.restype = i32
    L.fqb_last_path.argtypes = [vp, C.POINTER(u64 * 3)]
    L.fqb_last_path.restype = i32
    L.fqb_batch_begin.argtypes = [vp, u32]
    L.fqb_batch_begin.restype = i32
    L.fqb_batch_close.argtypes = [vp]
    L.fqb_batch_close.restype = i32
    L.fqb_next_batch.argtypes = [vp, C.POINTER(Batch)]
    L.fqb_next_batch.restype = i32
    L.fqb_release_batch.argtypes = [vp, u64]
    L.fqb_release_batch.restype = i32
    L.fqb_batch_cancel.argtypes = [vp]
    L.fqb_batch_cancel.restype = i32
    L.fqb_batch_end.argtypes = [vp, C.POINTER(Result)]
    L.fqb_batch_end.restype = i32
    L.fqb_comm_unique_id.argtypes = [vp]
    L.fqb_comm_unique_id.restype = i32
    L.fqb_comm_init.argtypes = [vp, i32, i32, vp]
    L.fqb_comm_init.restype = i32
    for n in ("fqb_comm_destroy", "fqb_comm_rank", "fqb_comm_world"):
        getattr(L, n).argtypes = [vp]
        getattr(L, n).restype = i32
    L.fqb_allreduce.argtypes = [vp, vp]
    L.fqb_allreduce.restype = i32
    L.fqb_fetch_reduced.argtypes = [vp, vp, vp, vp]
    L.fqb_fetch_reduced.restype = i32
    L.fqb_device_exchange.argtypes = [vp]
    L.fqb_device_exchange.restype = vp
    L.fqb_exchange_words.argtypes = [vp]
    L.fqb_exchange_words.restype = C.c_size_t
    L.fqb_filter_device.argtypes = [vp, vp, u64, vp, u64, u64, u32, vp, u64, vp]
    L.fqb_filter_device.restype = i32
    L.fqb_fetch_filter.argtypes = [vp, vp, C.POINTER(u64), C.POINTER(u64)]
    L.fqb_fetch_filter.restype = i32
    if L.fqb_abi_version() != ABI_VERSION:
        raise RuntimeError("libfastq_b200.so ABI version mismatch")
    _lib = L
    return L
