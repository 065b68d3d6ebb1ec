"""ctypes front end of the CPU oracle (oracle/fastq_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  Nothing under fastq_rs_b200/ imports this module.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libfastq_oracle.so")

OK, E_HEADER, E_SEP, E_LENGTH, E_TOO_LONG, E_TRUNCATED, E_IO = range(7)
BUFSIZE = 68 * 1024
SEED = 0xFA57A11CE5EED001

ERR_NAMES = {
    OK: "ok",
    E_HEADER: "header",
    E_SEP: "sep",
    E_LENGTH: "length",
    E_TOO_LONG: "too_long",
    E_TRUNCATED: "truncated",
    E_IO: "io",
}


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "fastq_oracle.c")
    hdr = os.path.join(_HERE, "fastq_oracle.h")
    stale = (not os.path.exists(_SO)) or any(
        os.path.exists(p) and os.path.getmtime(p) > os.path.getmtime(_SO) for p in (src, hdr)
    )
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B", "libfastq_oracle.so"],
                              stdout=subprocess.DEVNULL)
    return _SO


class _EachResult(C.Structure):
    _fields_ = [("status", C.c_int), ("finished", C.c_int), ("n_delivered", C.c_uint64),
                ("err_offset", C.c_uint64)]


class _Stats(C.Structure):
    _fields_ = [("max_len", C.c_uint32), ("n_records", C.c_uint64), ("n_bases", C.c_uint64),
                ("clip_seq", C.c_uint64), ("clip_qual", C.c_uint64),
                ("base_hist", C.POINTER(C.c_uint64)), ("qual_hist", C.POINTER(C.c_uint64)),
                ("len_hist", C.POINTER(C.c_uint64))]


class _RefRecord(C.Structure):
    _fields_ = [("data", C.POINTER(C.c_uint8)), ("len", C.c_size_t), ("head", C.c_size_t),
                ("seq", C.c_size_t), ("sep", C.c_size_t), ("qual", C.c_size_t)]


class _Reader(C.Structure):
    _fields_ = [("data", C.c_void_p), ("len", C.c_size_t), ("pos", C.c_size_t),
                ("max_read", C.c_size_t)]


class _RecordSet(C.Structure):
    _fields_ = [("buffer", C.POINTER(C.c_uint8)), ("bufsize", C.c_size_t),
                ("records", C.c_void_p), ("n_records", C.c_size_t), ("cap_records", C.c_size_t)]


_EACH_FN = C.CFUNCTYPE(C.c_int, C.c_void_p, C.POINTER(_RefRecord), C.c_uint64)

_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        L.fqo_each.argtypes = [C.POINTER(_Reader), C.c_size_t, _EACH_FN, C.c_void_p,
                               C.POINTER(_EachResult)]
        L.fqo_each.restype = None
        L.fqo_stats_new.argtypes = [C.c_uint32]
        L.fqo_stats_new.restype = C.POINTER(_Stats)
        L.fqo_stats_free.argtypes = [C.POINTER(_Stats)]
        L.fqo_stats_free.restype = None
        L.fqo_each_stats.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                     C.POINTER(_Stats), C.POINTER(_EachResult)]
        L.fqo_each_stats.restype = None
        L.fqo_each_index.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                     C.c_void_p, C.c_size_t, C.POINTER(_EachResult)]
        L.fqo_each_index.restype = None
        L.fqo_each_filter.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t, C.c_int, C.c_void_p,
                                      C.c_size_t, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64),
                                      C.POINTER(_EachResult)]
        L.fqo_each_filter.restype = None
        L.fqo_parallel_each_stats.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                              C.c_int, C.POINTER(_Stats), C.c_void_p]
        L.fqo_parallel_each_stats.restype = C.c_int
        L.fqo_parallel_each_count.argtypes = [C.c_void_p, C.c_size_t, C.c_size_t, C.c_size_t,
                                              C.c_int, C.POINTER(C.c_uint64)]
        L.fqo_parallel_each_count.restype = C.c_int
        L.fqo_synth_fixed.argtypes = [C.c_uint64, C.c_uint32, C.c_uint64, C.c_size_t, C.c_void_p]
        L.fqo_synth_fixed.restype = None
        L.fqo_synth_var.argtypes = [C.c_uint64, C.c_uint64, C.c_uint64, C.c_void_p]
        L.fqo_synth_var.restype = C.c_size_t
        L.fqo_synth_var_len.argtypes = [C.c_uint64, C.c_uint64]
        L.fqo_synth_var_len.restype = C.c_uint32
        L.fqo_splitmix64.argtypes = [C.c_uint64]
        L.fqo_splitmix64.restype = C.c_uint64
        for name in ("fqo_rec_head", "fqo_rec_seq", "fqo_rec_qual", "fqo_rec_sepline"):
            f = getattr(L, name)
            f.argtypes = [C.POINTER(_RefRecord), C.POINTER(C.POINTER(C.c_uint8)),
                          C.POINTER(C.c_size_t)]
            f.restype = None
        L.fqo_record_sets_new.argtypes = [C.POINTER(_Reader), C.c_size_t]
        L.fqo_record_sets_new.restype = C.c_void_p
        L.fqo_record_sets_next.argtypes = [C.c_void_p, C.POINTER(C.POINTER(_RecordSet))]
        L.fqo_record_sets_next.restype = C.c_int
        L.fqo_record_sets_free.argtypes = [C.c_void_p]
        L.fqo_record_sets_free.restype = None
        L.fqo_record_set_free.argtypes = [C.POINTER(_RecordSet)]
        L.fqo_record_set_free.restype = None
        L.fqo_record_set_get.argtypes = [C.POINTER(_RecordSet), C.c_size_t, C.POINTER(_RefRecord)]
        L.fqo_record_set_get.restype = None
        L.fqo_validate_dna.argtypes = [C.POINTER(_RefRecord)]
        L.fqo_validate_dnan.argtypes = [C.POINTER(_RefRecord)]
        _lib = L
    return _lib


def _as_u8(data) -> np.ndarray:
    if isinstance(data, np.ndarray):
        assert data.dtype == np.uint8
        return np.ascontiguousarray(data)
    return np.frombuffer(bytes(data), dtype=np.uint8)


@dataclass
class EachResult:
    status: int
    finished: bool
    n_records: int
    err_offset: int

    @property
    def error(self) -> str:
        return ERR_NAMES[self.status]


@dataclass
class Stats:
    max_len: int
    n_records: int
    n_bases: int
    clip_seq: int
    clip_qual: int
    base_hist: np.ndarray  # [P, 6] u64
    qual_hist: np.ndarray  # [P, 256] u64
    len_hist: np.ndarray   # [P + 2] u64


def _copy_stats(sp) -> Stats:
    s = sp.contents
    P = s.max_len
    return Stats(
        P, s.n_records, s.n_bases, s.clip_seq, s.clip_qual,
        np.ctypeslib.as_array(s.base_hist, shape=(P * 6,)).copy().reshape(P, 6),
        np.ctypeslib.as_array(s.qual_hist, shape=(P * 256,)).copy().reshape(P, 256),
        np.ctypeslib.as_array(s.len_hist, shape=(P + 2,)).copy(),
    )


@dataclass
class Rec:
    head: bytes
    seq: bytes
    qual: bytes
    sep: bytes
    raw: bytes
    offset: int
    valid_dna: bool
    valid_dnan: bool


def each(data, callback=None, bufsize: int = BUFSIZE, max_read: int = 0):
    """Parser::each over `data` (src/lib.rs:221).  Returns (EachResult, [Rec...]).  If
    `callback` is given it is called per record and may return False to stop."""
    L = lib()
    a = _as_u8(data)
    out = []

    def view(fn, rp):
        p = C.POINTER(C.c_uint8)()
        n = C.c_size_t()
        fn(rp, C.byref(p), C.byref(n))
        return C.string_at(p, n.value) if n.value else b""

    def cb(_user, rp, off):
        r = rp.contents
        rec = Rec(view(L.fqo_rec_head, rp), view(L.fqo_rec_seq, rp), view(L.fqo_rec_qual, rp),
                  view(L.fqo_rec_sepline, rp), C.string_at(r.data, r.len), off,
                  bool(L.fqo_validate_dna(rp)), bool(L.fqo_validate_dnan(rp)))
        out.append(rec)
        if callback is not None:
            return 1 if callback(rec) else 0
        return 1

    rd = _Reader(a.ctypes.data, a.size, 0, max_read)
    res = _EachResult()
    L.fqo_each(C.byref(rd), bufsize, _EACH_FN(cb), None, C.byref(res))
    return EachResult(res.status, bool(res.finished), res.n_delivered, res.err_offset), out


def record_sets(data, bufsize: int = BUFSIZE, max_read: int = 0):
    """Parser::record_sets (src/lib.rs:430).  Returns (status, [[Rec...] per set])."""
    L = lib()
    a = _as_u8(data)
    rd = _Reader(a.ctypes.data, a.size, 0, max_read)
    it = L.fqo_record_sets_new(C.byref(rd), bufsize)
    sets = []
    status = OK

    def view(fn, rp):
        p = C.POINTER(C.c_uint8)()
        n = C.c_size_t()
        fn(rp, C.byref(p), C.byref(n))
        return C.string_at(p, n.value) if n.value else b""

    while True:
        sp = C.POINTER(_RecordSet)()
        rc = L.fqo_record_sets_next(it, C.byref(sp))
        if rc == 0:
            break
        if rc < 0:
            status = -rc
            break
        recs = []
        for i in range(sp.contents.n_records):
            r = _RefRecord()
            L.fqo_record_set_get(sp, i, C.byref(r))
            rp = C.pointer(r)
            recs.append(Rec(view(L.fqo_rec_head, rp), view(L.fqo_rec_seq, rp),
                            view(L.fqo_rec_qual, rp), view(L.fqo_rec_sepline, rp),
                            C.string_at(r.data, r.len), -1,
                            bool(L.fqo_validate_dna(rp)), bool(L.fqo_validate_dnan(rp))))
        sets.append(recs)
        L.fqo_record_set_free(sp)
    L.fqo_record_sets_free(it)
    return status, sets


def each_stats(data, max_len: int, bufsize: int = BUFSIZE, max_read: int = 0):
    L = lib()
    a = _as_u8(data)
    sp = L.fqo_stats_new(max_len)
    res = _EachResult()
    L.fqo_each_stats(a.ctypes.data, a.size, bufsize, max_read, sp, C.byref(res))
    st = _copy_stats(sp)
    L.fqo_stats_free(sp)
    return EachResult(res.status, bool(res.finished), res.n_delivered, res.err_offset), st


def each_index(data, bufsize: int = BUFSIZE, max_read: int = 0, cap: int | None = None):
    """Returns (EachResult, idx[n,5] u64): start, head_nl, seq_nl, sep_nl, qual_nl (stream
    offsets) of every record each() delivers."""
    L = lib()
    a = _as_u8(data)
    if cap is None:
        cap = a.size // 4 + 1
    out = np.zeros((cap, 5), dtype=np.uint64)
    res = _EachResult()
    L.fqo_each_index(a.ctypes.data, a.size, bufsize, max_read, out.ctypes.data, cap,
                     C.byref(res))
    n = min(cap, res.n_delivered)
    return EachResult(res.status, bool(res.finished), res.n_delivered, res.err_offset), out[:n]


def each_filter(data, mode: int, bufsize: int = BUFSIZE, max_read: int = 0):
    """each() with a closure that writes the records passing validate_dna (mode 1) / validate_dnan
    (mode 2) / every record (mode 0) verbatim.  Returns (EachResult, n_kept, bytes written)."""
    L = lib()
    a = _as_u8(data)
    out = np.empty(max(a.size, 1), dtype=np.uint8)
    res = _EachResult()
    nk, nb = C.c_uint64(0), C.c_uint64(0)
    L.fqo_each_filter(a.ctypes.data, a.size, bufsize, max_read, mode, out.ctypes.data, out.size,
                      C.byref(nk), C.byref(nb), C.byref(res))
    return (EachResult(res.status, bool(res.finished), res.n_delivered, res.err_offset), nk.value,
            out[:nb.value].tobytes())


def parallel_each_stats(data, max_len: int, n_threads: int, bufsize: int = BUFSIZE,
                        max_read: int = 0):
    L = lib()
    a = _as_u8(data)
    sp = L.fqo_stats_new(max_len)
    sets = np.zeros(max(n_threads, 1), dtype=np.uint64)
    rc = L.fqo_parallel_each_stats(a.ctypes.data, a.size, bufsize, max_read, n_threads, sp,
                                   sets.ctypes.data)
    st = _copy_stats(sp)
    L.fqo_stats_free(sp)
    return rc, st, sets


def parallel_each_count(data, n_threads: int, bufsize: int = BUFSIZE, max_read: int = 0):
    L = lib()
    a = _as_u8(data)
    n = C.c_uint64(0)
    rc = L.fqo_parallel_each_count(a.ctypes.data, a.size, bufsize, max_read, n_threads,
                                   C.byref(n))
    return rc, n.value


def synth_fixed(n_bytes: int, L_read: int = 150, byte_off: int = 0, seed: int = SEED) -> np.ndarray:
    out = np.empty(n_bytes, dtype=np.uint8)
    lib().fqo_synth_fixed(seed, L_read, byte_off, n_bytes, out.ctypes.data)
    return out


def synth_fixed_records(n_records: int, L_read: int = 150, first: int = 0, seed: int = SEED):
    rb = 17 + 2 * (L_read + 1) + 2
    return synth_fixed(n_records * rb, L_read, first * rb, seed)


def synth_var(n_records: int, first: int = 0, seed: int = SEED) -> np.ndarray:
    L = lib()
    n = L.fqo_synth_var(seed, first, n_records, None)
    out = np.empty(n, dtype=np.uint8)
    L.fqo_synth_var(seed, first, n_records, out.ctypes.data)
    return out


def synth_var_len(rec: int, seed: int = SEED) -> int:
    return lib().fqo_synth_var_len(seed, rec)
