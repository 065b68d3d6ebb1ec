"""Engine: thin Python owner of one fqb_ctx (one per GPU / per process).

torch is used for device memory and stream handles only; all compute goes through the C ABI
(include/fastq_b200.h) into the sm_100a kernels in csrc/.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib
from ._lib import (E_HEADER, E_LENGTH, E_PHASE, E_SEP, E_TOO_LONG, E_TRUNCATED, F_EOF, F_FRONT16, F_HIST,
                   F_INDEX, F_INFER_START, F_LINE_START, F_PARTIAL, OK)


class FastqError(ValueError):
    """Mirror of the reference's io::Error(InvalidData, msg) for grammar errors
    (src/records.rs:143-146,157-160,233-238; src/lib.rs:278-291)."""

    def __init__(self, status: int, offset: int, n_delivered: int):
        self.status = status
        self.offset = offset
        self.n_delivered = n_delivered
        super().__init__(_lib.lib().fqb_strerror(status).decode())

    @property
    def kind(self) -> str:
        return "InvalidData" if 1 <= self.status <= 5 else "Other"


@dataclass
class Outcome:
    """What Parser::each returns (src/lib.rs:221-238) plus where it stopped."""
    status: int
    finished: bool
    n_records: int
    n_lines: int
    err_offset: int
    tail_offset: int | None
    line_phase: int = 0

    def raise_for_status(self):
        if self.status != OK:
            raise FastqError(self.status, self.err_offset, self.n_records)


class Stats:
    """View of one statistics block (layout in include/fastq_b200.h)."""

    def __init__(self, max_len: int, words: np.ndarray):
        L = _lib.lib()
        self.max_len = P = max_len
        self.words = words
        lo, bo, qo = (L.fqb_stats_len_hist_off(P), L.fqb_stats_base_hist_off(P),
                      L.fqb_stats_qual_hist_off(P))
        self.len_hist = words[lo:lo + P + 2]
        self.base_hist = words[bo:bo + 6 * P].reshape(P, 6)
        self.qual_hist = words[qo:qo + 256 * P].reshape(P, 256)

    n_records = property(lambda s: int(s.words[0]))
    n_bases = property(lambda s: int(s.words[1]))
    clip_seq = property(lambda s: int(s.words[2]))
    clip_qual = property(lambda s: int(s.words[3]))


def _check(ctx, rc: int, what: str):
    if rc != OK:
        L = _lib.lib()
        detail = L.fqb_last_error(ctx).decode() if ctx else ""
        raise RuntimeError(f"{what} failed: {L.fqb_strerror(rc).decode()} ({rc}) {detail}")


class Engine:
    def __init__(self, max_len: int = 150, device: int = 0, slot_bytes: int = 0, n_slots: int = 0):
        self.L = _lib.lib()
        self.max_len = max_len
        self.device = device
        cfg = _lib.Config(_lib.ABI_VERSION, device, max_len, 0, slot_bytes, n_slots, 0)
        h = C.c_void_p()
        rc = self.L.fqb_create(C.byref(cfg), C.byref(h))
        if rc != OK:
            raise RuntimeError(
                f"fqb_create failed ({rc}: {self.L.fqb_strerror(rc).decode()}); the FASTQ hot path "
                "needs a CUDA device -- there is no CPU fallback")
        self.ctx = h
        self.n_words = self.L.fqb_stats_words(max_len)

    def close(self):
        if getattr(self, "ctx", None):
            self.L.fqb_destroy(self.ctx)
            self.ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- in-HBM path ---------------------------------------------------------------------
    @staticmethod
    def _stream_ptr(stream):
        if stream is None:
            import torch
            stream = torch.cuda.current_stream()
        return C.c_void_p(getattr(stream, "cuda_stream", stream))

    def parse_device(self, d_bytes, n_own: int | None = None, n_avail: int | None = None, *,
                     hist: bool = True, index=None, line_base: int = 0, stream_offset: int = 0,
                     line_start: bool = True, eof: bool = True, front16: bool = False,
                     infer_start: bool = False, stream=None) -> None:
        """Enqueue delimit (+index) (+histograms) over a uint8 CUDA tensor (asynchronous).
        `d_bytes` may be a tensor or a raw device address."""
        ptr = d_bytes if isinstance(d_bytes, int) else d_bytes.data_ptr()
        total = None if isinstance(d_bytes, int) else d_bytes.numel()
        if n_avail is None:
            n_avail = total if n_own is None else max(n_own, total if total is not None else n_own)
        if n_own is None:
            n_own = n_avail
        flags = (F_HIST if hist else 0) | (F_INDEX if index is not None else 0)
        flags |= (F_LINE_START if line_start else 0) | (F_EOF if eof else 0) | (F_FRONT16 if front16 else 0)
        flags |= F_INFER_START if infer_start else 0
        sh = _lib.Shard(ptr, n_own, n_avail, stream_offset, line_base, flags, 0,
                        index.data_ptr() if index is not None else None,
                        index.numel() if index is not None else 0)
        _check(self.ctx, self.L.fqb_parse_device(self.ctx, C.byref(sh), self._stream_ptr(stream)),
               "fqb_parse_device")

    def fetch(self, want_stats: bool = True, stream=None):
        """Synchronise and copy out (Outcome, Stats | None)."""
        res = _lib.Result()
        words = np.zeros(self.n_words, dtype=np.uint64) if want_stats else None
        _check(self.ctx, self.L.fqb_fetch(self.ctx, self._stream_ptr(stream), C.byref(res),
                                          words.ctypes.data if want_stats else None), "fqb_fetch")
        return self._outcome(res), (Stats(self.max_len, words) if want_stats else None)

    @staticmethod
    def _outcome(res) -> Outcome:
        tail = None if res.tail_offset == _lib.NO_OFFSET else int(res.tail_offset)
        return Outcome(int(res.status), bool(res.finished), int(res.n_records), int(res.n_lines),
                       int(res.err_offset), tail, int(res.line_phase))

    def _device_view(self, ptr: int, n_words: int):
        import torch

        class _Arr:
            __cuda_array_interface__ = {"shape": (n_words,), "typestr": "<i8", "data": (ptr, False), "version": 3}
        return torch.as_tensor(_Arr(), device=f"cuda:{self.device}")

    def device_stats(self):
        """The device-resident stats block of the last parse as an int64 CUDA tensor view
        (same bits as u64; for an all_reduce)."""
        if getattr(self, "_stats_view", None) is None:
            self._stats_view = self._device_view(self.L.fqb_device_stats(self.ctx), self.n_words)
        return self._stats_view

    def device_result(self):
        """The outcome of the last parse_device as 8 device-resident int64 words
        [status, finished, n_records, n_lines, err_offset, tail_offset, line_phase, 0] -- valid in
        stream order after the parse; lets the N-rank driver all-gather outcomes without a host sync."""
        if getattr(self, "_result_view", None) is None:
            self._result_view = self._device_view(self.L.fqb_device_result(self.ctx), 8)
        return self._result_view

    # ---- N ranks: the one collective, behind the ABI (NCCL bound by the library itself) ------
    def comm_unique_id(self) -> bytes:
        buf = (C.c_uint8 * _lib.COMM_ID_BYTES)()
        _check(self.ctx, self.L.fqb_comm_unique_id(buf), "fqb_comm_unique_id")
        return bytes(buf)

    def comm_init(self, rank: int, world: int, uid: bytes | None = None) -> None:
        """collective over all ranks; `uid` = rank 0's comm_unique_id(), handed around by the caller"""
        buf = (C.c_uint8 * _lib.COMM_ID_BYTES).from_buffer_copy(uid) if uid is not None else None
        _check(self.ctx, self.L.fqb_comm_init(self.ctx, rank, world, buf), "fqb_comm_init")
        self._xchg_view = None

    def comm_world(self) -> int:
        return int(self.L.fqb_comm_world(self.ctx))

    def allreduce(self, stream=None) -> None:
        """enqueue the all-reduce(sum, u64) of [statistics block | world x 8 outcome words] of the last parse"""
        _check(self.ctx, self.L.fqb_allreduce(self.ctx, self._stream_ptr(stream)), "fqb_allreduce")

    def fetch_reduced(self, want_stats: bool = True, stream=None):
        """wait for allreduce(): (global stats words | None, outcomes[world, 8]) -- one copy, one synchronisation"""
        world = self.comm_world()
        words = np.zeros(self.n_words, dtype=np.uint64) if want_stats else None
        outs = np.zeros(8 * world, dtype=np.uint64)
        _check(self.ctx, self.L.fqb_fetch_reduced(self.ctx, self._stream_ptr(stream),
                                                  words.ctypes.data if want_stats else None, outs.ctypes.data),
               "fqb_fetch_reduced")
        return words, outs.view(np.int64).reshape(world, 8)

    def device_exchange(self):
        """[statistics block | world x 8 outcome words] of the last parse as an int64 CUDA tensor view -- the
        send buffer of the one collective, for callers that run it with their own library"""
        if getattr(self, "_xchg_view", None) is None:
            self._xchg_view = self._device_view(self.L.fqb_device_exchange(self.ctx), int(self.L.fqb_exchange_words(self.ctx)))
        return self._xchg_view

    def count_lines(self, d_bytes, n: int | None = None, stream=None) -> int:
        n = d_bytes.numel() if n is None else n
        sp = self._stream_ptr(stream)
        _check(self.ctx, self.L.fqb_count_lines_device(self.ctx, d_bytes.data_ptr(), n, sp),
               "fqb_count_lines_device")
        out = C.c_uint64()
        _check(self.ctx, self.L.fqb_fetch_line_count(self.ctx, sp, C.byref(out)), "fqb_fetch_line_count")
        return out.value

    def count_lines_async(self, d_bytes, n: int | None = None, stream=None) -> None:
        n = d_bytes.numel() if n is None else n
        _check(self.ctx, self.L.fqb_count_lines_device(self.ctx, d_bytes.data_ptr(), n,
                                                       self._stream_ptr(stream)),
               "fqb_count_lines_device")

    # ---- record filter (validate_dna / validate_dnan + Record::write) ------------------------
    def filter_device(self, d_bytes, index, n_records: int, mode: int, out, *, stream_offset: int = 0,
                      first_offset: int | None = None, stream=None) -> None:
        """Enqueue: write the records of a parsed shard whose seq() passes the predicate (mode:
        _lib.KEEP_ALL / KEEP_DNA / KEEP_DNAN, src/records.rs:19-33) verbatim into the uint8 CUDA tensor
        `out`, densely, in order.  `index` = the int32 line-end index parse_device filled."""
        first = stream_offset if first_offset is None else first_offset
        _check(self.ctx, self.L.fqb_filter_device(
            self.ctx, d_bytes.data_ptr(), stream_offset, index.data_ptr(), n_records, first, mode,
            out.data_ptr() if out is not None else None, out.numel() if out is not None else 0,
            self._stream_ptr(stream)), "fqb_filter_device")

    def fetch_filter(self, stream=None) -> tuple[int, int]:
        """(records kept, bytes they occupy) of the last filter_device (waits for it)."""
        nk, nb = C.c_uint64(), C.c_uint64()
        _check(self.ctx, self.L.fqb_fetch_filter(self.ctx, self._stream_ptr(stream), C.byref(nk), C.byref(nb)),
               "fqb_fetch_filter")
        return nk.value, nb.value

    def last_path(self) -> dict:
        """Which path produced the result of the last fetch(): {"exact": the speculative pass was abandoned,
        "retried": a record in the middle of the shard was bad and the bytes in front of it were parsed a second time
        (speculatively), "predicted": windows predicted, "scanned": windows scanned} (diagnostics)."""
        out = (C.c_uint64 * 3)()
        _check(self.ctx, self.L.fqb_last_path(self.ctx, C.byref(out)), "fqb_last_path")
        return {"exact": bool(out[0] & 1), "retried": bool(out[0] & 2), "predicted": int(out[1]), "scanned": int(out[2])}

    def last_scan_ms(self) -> float:
        return float(self.L.fqb_last_scan_ms(self.ctx))

    def last_index_ms(self) -> float:
        return float(self.L.fqb_last_index_ms(self.ctx))

    def launch_count(self) -> int:
        return int(self.L.fqb_launch_count(self.ctx))

    # ---- host path -------------------------------------------------------------------------
    def parse_host(self, data, *, hist: bool = True, want_index: bool = False, want_stats: bool = True,
                   partial: bool = False, stream_offset: int = 0):
        """Delimit (+histograms) host bytes end to end through the pinned ring.
        `data`: bytes-like / uint8 ndarray / (address, nbytes).  Returns (Outcome, Stats|None, index|None)
        where index = u64 stream offsets of every line end.  `partial`: one refill of a longer stream
        (FQB_F_PARTIAL): an incomplete trailing record is reported in Outcome.tail_offset, not as an error.
        `stream_offset`: stream offset of data[0] (all offsets that come back are stream offsets)."""
        if isinstance(data, tuple):
            addr, n = data
            keep = None
        else:
            keep = np.frombuffer(data, dtype=np.uint8) if not isinstance(data, np.ndarray) else \
                np.ascontiguousarray(data.view(np.uint8))
            addr, n = (keep.ctypes.data if keep.size else 0), keep.size
        res = _lib.Result()
        words = np.zeros(self.n_words, dtype=np.uint64) if want_stats else None
        idx = np.empty(max(n, 1), dtype=np.uint32) if want_index else None
        n_idx = C.c_uint64(0)
        flags = (F_HIST if hist else 0) | (F_INDEX if want_index else 0) | (F_PARTIAL if partial else 0)
        _check(self.ctx, self.L.fqb_parse_host(
            self.ctx, addr, n, stream_offset, flags, C.byref(res), words.ctypes.data if want_stats else None,
            idx.ctypes.data if want_index else None, idx.size if want_index else 0, C.byref(n_idx)),
            "fqb_parse_host")
        index = expand_index(idx[:n_idx.value], stream_offset) if want_index else None
        return self._outcome(res), (Stats(self.max_len, words) if want_stats else None), index

    # ---- streaming ring (thread_reader protocol) -------------------------------------------
    def stream_begin(self, hist: bool = True):
        _check(self.ctx, self.L.fqb_stream_begin(self.ctx, F_HIST if hist else 0), "fqb_stream_begin")

    def stream_acquire(self):
        p, cap = C.c_void_p(), C.c_uint64()
        _check(self.ctx, self.L.fqb_stream_acquire(self.ctx, C.byref(p), C.byref(cap)), "fqb_stream_acquire")
        return (C.c_uint8 * cap.value).from_address(p.value)

    def stream_submit(self, n_valid: int):
        _check(self.ctx, self.L.fqb_stream_submit(self.ctx, n_valid), "fqb_stream_submit")

    def stream_finish(self, want_stats: bool = True):
        res = _lib.Result()
        words = np.zeros(self.n_words, dtype=np.uint64) if want_stats else None
        _check(self.ctx, self.L.fqb_stream_finish(self.ctx, C.byref(res),
                                                  words.ctypes.data if want_stats else None),
               "fqb_stream_finish")
        return self._outcome(res), (Stats(self.max_len, words) if want_stats else None)

    # ---- batch mode: the generic-closure path, asynchronous ------------------------------------
    def batch_begin(self, hist: bool = False) -> bool:
        """False if the context is busy with another stream (one ring per context) or has too few slots"""
        rc = self.L.fqb_batch_begin(self.ctx, F_HIST if hist else 0)
        if rc == _lib.E_STATE:
            return False
        _check(self.ctx, rc, "fqb_batch_begin")
        return True

    def batch_acquire(self):
        """producer: the next pinned slot (None once the consumer has cancelled)"""
        p, cap = C.c_void_p(), C.c_uint64()
        rc = self.L.fqb_stream_acquire(self.ctx, C.byref(p), C.byref(cap))
        if rc == _lib.E_CANCELLED:
            return None
        _check(self.ctx, rc, "fqb_stream_acquire")
        return (C.c_uint8 * cap.value).from_address(p.value)

    def batch_close(self):
        _check(self.ctx, self.L.fqb_batch_close(self.ctx), "fqb_batch_close")

    def next_batch(self):
        """consumer: (data: uint8 view of the pinned bytes -- the records, then whatever of the stream follows them in
        the ring --, ends: int64 [n, 4] offsets of the line ends within data, status, err_offset, stream_offset,
        first_record, last, token) -- valid until release_batch(token)"""
        b = _lib.Batch()
        rc = self.L.fqb_next_batch(self.ctx, C.byref(b))
        if rc == _lib.E_CANCELLED:
            return None
        _check(self.ctx, rc, "fqb_next_batch")
        n = int(b.n_records)
        data = np.ctypeslib.as_array((C.c_uint8 * int(b.n_avail)).from_address(b.bytes)) if b.n_avail else np.empty(0, np.uint8)
        if n:
            raw = np.ctypeslib.as_array((C.c_uint32 * (4 * n)).from_address(b.line_ends))
            ends = (raw - np.uint32(b.stream_offset & 0xFFFFFFFF)).astype(np.uint32).astype(np.int64).reshape(n, 4)
        else:
            ends = np.empty((0, 4), dtype=np.int64)
        return data, ends, int(b.status), int(b.err_offset), int(b.stream_offset), int(b.first_record), bool(b.last), int(b.token)

    def release_batch(self, token: int):
        _check(self.ctx, self.L.fqb_release_batch(self.ctx, token), "fqb_release_batch")

    def batch_cancel(self):
        self.L.fqb_batch_cancel(self.ctx)

    def batch_end(self) -> Outcome:
        res = _lib.Result()
        _check(self.ctx, self.L.fqb_batch_end(self.ctx, C.byref(res)), "fqb_batch_end")
        return self._outcome(res)

    # ---- synthetic data (bench / tests) ----------------------------------------------------
    def synth_fixed(self, out, n: int, byte_off: int = 0, read_len: int = 150,
                    seed: int = _lib.SYNTH_SEED, stream=None):
        ptr = out if isinstance(out, int) else out.data_ptr()
        _check(self.ctx, self.L.fqb_synth_fixed_device(ptr, n, byte_off, read_len, seed,
                                                       self._stream_ptr(stream)), "fqb_synth_fixed_device")

    def synth_var(self, n_records: int, first: int = 0, seed: int = _lib.SYNTH_SEED, pad: int = 0):
        """Variable-length (50..300 bp) records [first, first+n) as a uint8 CUDA tensor."""
        import torch
        dev = f"cuda:{self.device}"
        sizes = torch.empty(n_records, dtype=torch.int64, device=dev)
        sp = self._stream_ptr(None)
        _check(self.ctx, self.L.fqb_synth_var_sizes_device(sizes.data_ptr(), first, n_records, seed, sp),
               "fqb_synth_var_sizes_device")
        offs = torch.zeros(n_records + 1, dtype=torch.int64, device=dev)
        torch.cumsum(sizes, 0, out=offs[1:])
        total = int(offs[-1].item())
        out = torch.zeros(total + pad, dtype=torch.uint8, device=dev)
        _check(self.ctx, self.L.fqb_synth_var_device(out.data_ptr(), offs.data_ptr(), first, n_records,
                                                     seed, sp), "fqb_synth_var_device")
        return out, total


def expand_index(lo32: np.ndarray, base: int = 0) -> np.ndarray:
    """Low-32-bit line-end offsets -> u64 stream offsets (offsets are strictly increasing, so a
    decrease marks a 4 GiB wrap).  `base`: stream offset the first entries lie behind (< 4 GiB before them)."""
    if lo32.size == 0:
        return lo32.astype(np.uint64)
    rel = (lo32 - np.uint32(base & 0xFFFFFFFF)).astype(np.uint32)      # offsets relative to base, mod 2^32
    lo = rel.astype(np.uint64)
    wraps = np.zeros(lo.size, dtype=np.uint64)
    wraps[1:] = np.cumsum(rel[1:] < rel[:-1]).astype(np.uint64)
    return lo + (wraps << np.uint64(32)) + np.uint64(base)
