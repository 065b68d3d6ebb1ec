"""Host-side mirror of the `fastq` crate's interface for the delimiting path.

Same names, argument meaning and error behaviour as the reference (aseyboldt/fastq-rs 0.6.0):

    Parser(reader).each(func) -> bool                      src/lib.rs:221-238
    Parser(reader).parallel_each(n_threads, func) -> list  src/lib.rs:509-566
    Parser(reader).ref_iter() -> RecordRefIter             src/lib.rs:208, 241-304
    RecordSet.iter()/len()/is_empty()                      src/lib.rs:306-336
    Record: head()/seq()/qual()/write()/validate_dna(n)()  src/records.rs:5-34
    RefRecord / OwnedRecord / to_owned_record              src/records.rs:36-54,93-129,165-175
    each_zipped(parser1, parser2, callback)                src/lib.rs:577-609

plus the fast paths that never materialise records on the host: Parser.count(), Parser.stats().

Record delimiting ('\n' scan, '@'/'+'/length validation, record offsets) always runs on the GPU
through the C ABI; closures run here over the returned index, borrowing the caller's bytes.

Readers are consumed in refills of `chunk_bytes` (bounded memory, like the reference's Buffer --
src/buffer.rs:51-100 -- at a GPU-sized granularity): each refill is delimited with FQB_F_PARTIAL,
and the incomplete record at its end is carried over in front of the next one (Buffer::clean).
"""
from __future__ import annotations

import queue
import threading
from typing import Callable, Iterator, Optional

import numpy as np

from .engine import Engine, FastqError, Outcome, Stats

BUFSIZE = 68 * 1024  # src/lib.rs:129
CHUNK_BYTES = 32 << 20  # refill size of the generic-closure path (the reference refills 68 KiB at a time)

_default_engines: dict = {}


def default_engine(max_len: int = 150, device: int = 0, k: int = 0) -> Engine:
    """the k-th shared engine for (max_len, device): a context serves one stream at a time, so parsers that run
    side by side (each_zipped) each take their own"""
    key = (max_len, device, k)
    if key not in _default_engines:
        _default_engines[key] = Engine(max_len=max_len, device=device, slot_bytes=8 << 20, n_slots=4)
    return _default_engines[key]


def _trim_winline(b: bytes) -> bytes:  # src/records.rs:65-73
    return b[:-1] if b.endswith(b"\r") else b


class OwnedRecord:  # src/records.rs:49-54, 99-129
    def __init__(self, head: bytes, seq: bytes, sep: Optional[bytes], qual: bytes):
        self._head, self._seq, self.sep, self._qual = head, seq, sep, qual

    def head(self) -> bytes:
        return self._head

    def seq(self) -> bytes:
        return self._seq

    def qual(self) -> bytes:
        return self._qual

    def write(self, writer) -> int:
        parts = [b"@", self._head, b"\n", self._seq, b"\n", self.sep if self.sep is not None else b"+",
                 b"\n", self._qual, b"\n"]
        n = 0
        for p in parts:
            writer.write(p)
            n += len(p)
        return n

    def validate_dna(self) -> bool:
        return all(c in b"ACTG" for c in self._seq)

    def validate_dnan(self) -> bool:
        return all(c in b"ACTGN" for c in self._seq)

    def __eq__(self, o):
        return isinstance(o, OwnedRecord) and (self._head, self._seq, self.sep, self._qual) == \
            (o._head, o._seq, o.sep, o._qual)


class RefRecord:
    """Borrowed view (src/records.rs:36-45): `data` is the record's bytes, head/seq/sep/qual the
    record-relative offsets of its four line ends."""
    __slots__ = ("data", "_head", "_seq", "_sep", "_qual", "offset")

    def __init__(self, data: memoryview, head: int, seq: int, sep: int, qual: int, offset: int):
        self.data, self._head, self._seq, self._sep, self._qual, self.offset = data, head, seq, sep, qual, offset

    def head(self) -> bytes:  # skips '@'
        return _trim_winline(bytes(self.data[1:self._head]))

    def seq(self) -> bytes:
        return _trim_winline(bytes(self.data[self._head + 1:self._seq]))

    def qual(self) -> bytes:
        return _trim_winline(bytes(self.data[self._sep + 1:self._qual]))

    def write(self, writer) -> int:
        writer.write(bytes(self.data))
        return len(self.data)

    def validate_dna(self) -> bool:
        return all(c in b"ACTG" for c in self.seq())

    def validate_dnan(self) -> bool:
        return all(c in b"ACTGN" for c in self.seq())

    def to_owned_record(self) -> OwnedRecord:  # src/records.rs:165-175
        return OwnedRecord(self.head(), self.seq(),
                           _trim_winline(bytes(self.data[self._seq + 1:self._sep])), self.qual())


class RecordSet:  # src/lib.rs:306-336
    def __init__(self, buf: memoryview, starts: np.ndarray, ends4: np.ndarray):
        self._buf, self._starts, self._ends = buf, starts, ends4

    def iter(self) -> Iterator[RefRecord]:
        for i in range(len(self._starts)):
            yield _make_record(self._buf, int(self._starts[i]), self._ends[i])

    __iter__ = iter

    def len(self) -> int:
        return len(self._starts)

    __len__ = len

    def is_empty(self) -> bool:
        return len(self._starts) == 0


def _next_fill(fill_end: int, cur: int) -> int:
    """Buffer::replace_buffer + read_into (src/buffer.rs:30-48,74-100) in stream coordinates: the n = fill_end - cur
    leftover bytes are parked so that they end at ceil16(n); the read asks for n_free rounded down to 4096 (all of
    n_free below 4096).  Returns the new fill end (== fill_end when the buffer is full: "record too long")."""
    n = fill_end - cur
    new_end = (n + 15) & ~15
    free = BUFSIZE - new_end
    return fill_end + (free if free < 4096 else free - free % 4096)


def _owned_set(pend: list) -> "RecordSet":
    """a RecordSet that owns a copy of its records' bytes (src/lib.rs:308-311)"""
    if not pend:
        return RecordSet(memoryview(b""), np.empty(0, np.int64), np.empty((0, 4), np.int64))
    buf = b"".join(p[0] for p in pend)
    starts = np.empty(len(pend), np.int64)
    ends = np.empty((len(pend), 4), np.int64)
    at = 0
    for k, (raw, s_abs, e4) in enumerate(pend):
        starts[k] = at
        ends[k] = np.asarray(e4, dtype=np.int64) - s_abs + at
        at += len(raw)
    return RecordSet(memoryview(buf), starts, ends)


def _detect_point(d: "_Delimited", status: int, e: int):
    """Stream offset of the byte whose presence in the buffer makes from_buffer return the error `status` for the
    record at stream offset e (src/records.rs:201-247): '@' -> e itself; '+' -> the byte behind the sequence line;
    length mismatch -> the fourth line end.  None if the host's copy of the stream does not reach that far."""
    if status == 1:
        return e
    buf = d.buf
    rel = e - d.base
    pos = rel
    need = 2 if status == 2 else 4
    for k in range(need):
        nxt = bytes(buf[pos:pos + 70000]).find(b"\n")
        if nxt < 0:
            return None
        pos += nxt + 1
    return d.base + (pos if status == 2 else pos - 1)


def _make_record(buf: memoryview, start: int, ends) -> RefRecord:
    e0, e1, e2, e3 = (int(x) for x in ends)
    return RefRecord(buf[start:e3 + 1], e0 - start, e1 - start, e2 - start, e3 - start, start)


def _read_all(reader) -> np.ndarray:
    if isinstance(reader, np.ndarray):
        return np.ascontiguousarray(reader.view(np.uint8))
    if isinstance(reader, (bytes, bytearray, memoryview)):
        return np.frombuffer(bytes(reader), dtype=np.uint8)
    chunks = []
    while True:
        b = reader.read(1 << 22)
        if not b:
            break
        chunks.append(b)
    return np.frombuffer(b"".join(chunks), dtype=np.uint8)


def _fill(reader, view: np.ndarray) -> int:
    """Read until `view` is full or the reader is exhausted (short reads are topped up, the loop of
    src/buffer.rs:74-100); returns the bytes read."""
    got = 0
    mv = memoryview(view)
    while got < len(mv):
        if hasattr(reader, "readinto"):
            n = reader.readinto(mv[got:])
        else:
            b = reader.read(len(mv) - got)
            n = len(b)
            mv[got:got + n] = b
        if not n:
            break
        got += n
    return got


class _Delimited:
    """Result of the GPU delimiting pass over one refill / one batch of the stream: record starts and line ends
    (relative to `data`); `base` = stream offset of data[0], `delivered` = records of earlier refills.  `data`
    may reach beyond the last delivered record (the bad or incomplete record that follows it)."""

    def __init__(self, data: np.ndarray, outcome: Outcome, index: np.ndarray, base: int = 0, delivered: int = 0,
                 rel_ends: Optional[np.ndarray] = None, at_eof: bool = False):
        self.base, self.delivered, self.at_eof = base, delivered, at_eof
        n = outcome.n_records
        ends = rel_ends if rel_ends is not None else index[:4 * n].astype(np.int64).reshape(n, 4) - base   # relative to data[0]
        starts = np.empty(n, dtype=np.int64)
        if n:
            starts[0] = 0
            starts[1:] = ends[:-1, 3] + 1
        self.buf = memoryview(data)
        self.outcome, self.starts, self.ends = outcome, starts, ends

    def raise_for_status(self) -> None:
        o = self.outcome
        if o.status != 0:
            raise FastqError(o.status, o.err_offset, self.delivered + o.n_records)   # (stream offset)


class _SyncChannel:
    """std::sync::mpsc::sync_channel(bound) as parallel_each uses it (src/lib.rs:522): a bounded queue whose
    send() blocks while it is full and FAILS once the receiver is gone; the receiver's iterator ends when
    the sender is dropped and the queue is empty."""
    _END = object()

    def __init__(self, bound: int):
        self._q: queue.Queue = queue.Queue(maxsize=bound)
        self._rx_gone = threading.Event()
        self._closed = False

    def send(self, item) -> bool:
        while not self._rx_gone.is_set():
            try:
                self._q.put(item, timeout=0.02)
                return True
            except queue.Full:
                continue
        return False

    def close(self) -> None:            # drop(tx)
        if not self._closed:
            self._closed = True
            self.send(self._END)

    def receive(self):
        while True:
            item = self._q.get()
            if item is self._END:
                return
            yield item

    def hang_up(self) -> None:          # drop(rx)
        self._rx_gone.set()
        try:
            while True:
                self._q.get_nowait()
        except queue.Empty:
            pass


class RecordRefIter:  # src/lib.rs:241-304
    def __init__(self, chunks: Iterator[_Delimited]):
        self._chunks, self._d, self._i, self._cur, self._done = iter(chunks), None, -1, None, False

    def advance(self) -> None:
        self._cur = None
        while not self._done:
            if self._d is not None:
                self._i += 1
                if self._i < len(self._d.starts):
                    self._cur = _make_record(self._d.buf, int(self._d.starts[self._i]), self._d.ends[self._i])
                    return
                if self._d.outcome.status != 0:
                    self._done = True
                    self._d.raise_for_status()  # Err only after every earlier record was handed out
            self._d, self._i = next(self._chunks, None), -1   # refill (src/lib.rs:262-294)
            self._done = self._d is None

    def get(self) -> Optional[RefRecord]:
        return self._cur

    def close(self) -> None:
        """drop the parser: stops the reader thread of the batch mode, gives the ring back"""
        self._done = True
        close = getattr(self._chunks, "close", None)
        if close is not None:
            close()


class Parser:
    """Parser::new(reader) (src/lib.rs:200).  `reader`: bytes-like, uint8 ndarray, or an object
    with .read().  `max_len` = positions tracked by stats()."""

    def __init__(self, reader, engine: Optional[Engine] = None, max_len: int = 150,
                 chunk_bytes: Optional[int] = None):
        """`chunk_bytes`: None = readers are consumed through the asynchronous batch mode (ring slots of the
        engine); a number = synchronous refills of that size (FQB_F_PARTIAL + carry-over)."""
        self._reader = reader
        self._own_engine = engine is None
        self._max_len = max_len
        self._engine = engine or default_engine(max_len)
        self._chunk_bytes = None if chunk_bytes is None else max(1, int(chunk_bytes))

    def _chunks(self) -> Iterator[_Delimited]:
        """Delimit the stream refill by refill.  In-memory inputs are one refill; a reader is consumed
        `chunk_bytes` at a time with the incomplete record at the end of a refill carried over in front
        of the next (Buffer::clean, src/buffer.rs:51-72), so memory stays bounded by chunk_bytes + one
        record.  A refill that reports an error is the last one."""
        r, eng = self._reader, self._engine
        if isinstance(r, (bytes, bytearray, memoryview, np.ndarray)):
            data = _read_all(r)
            outcome, _, index = eng.parse_host(data, hist=False, want_index=True, want_stats=False)
            yield _Delimited(data, outcome, index, at_eof=True)
            return
        if self._chunk_bytes is None and hasattr(eng, "batch_begin"):
            k = 0
            while not eng.batch_begin():
                # the ring of this context is busy with another parser's stream (each_zipped): take the next of the
                # shared engines; an engine the caller passed in is the caller's to keep free
                if not self._own_engine:
                    raise RuntimeError("the engine is busy with another stream: give every concurrent Parser its own Engine")
                k += 1
                eng = self._engine = default_engine(self._max_len, eng.device, k)
            yield from self._batches()
            return
        # (synchronous refills: asked for, or the engine's ring is busy with another stream -- each_zipped over
        # two parsers that share an engine --, or a stand-in engine without the batch mode)
        left = np.empty(0, dtype=np.uint8)
        base = delivered = 0
        ch = self._chunk_bytes or CHUNK_BYTES
        while True:
            buf = np.empty(left.size + ch, dtype=np.uint8)
            buf[:left.size] = left
            n = _fill(r, buf[left.size:])
            eof = n < ch
            buf = buf[:left.size + n]
            outcome, _, index = eng.parse_host(buf, hist=False, want_index=True, want_stats=False,
                                               partial=not eof, stream_offset=base)
            yield _Delimited(buf, outcome, index, base, delivered, at_eof=eof)
            if eof or outcome.status != 0:
                return
            t = buf.size if outcome.tail_offset is None else outcome.tail_offset - base
            left, base, delivered = buf[t:], base + t, delivered + outcome.n_records

    def _batches(self) -> Iterator[_Delimited]:
        """The asynchronous form of _chunks (batch mode of the C ABI): a reader thread keeps the pinned ring full
        (thread_reader's protocol, src/thread_reader.rs:40-50) and the GPU keeps delimiting while this thread hands
        out the chunks already delimited.  The records of a batch borrow the pinned ring: they are valid until the
        next batch is asked for -- the lifetime of a RefRecord inside each() (src/lib.rs:248-252)."""
        eng, r = self._engine, self._reader      # (the caller has begun the batch mode)
        err: list = []

        def pump():
            try:
                while True:
                    slot = eng.batch_acquire()                   # blocks until a pinned slot is free
                    if slot is None:
                        return                                   # the consumer stopped
                    mv = memoryview(slot).cast("B")
                    if hasattr(r, "readinto"):
                        n = r.readinto(mv) or 0
                    else:
                        b = r.read(len(mv))
                        n = len(b)
                        mv[:n] = b
                    eng.stream_submit(n)
                    if not n:
                        eng.batch_close()
                        return
            except BaseException as e:                           # reader errors travel to the caller
                err.append(e)
                eng.batch_cancel()

        t = threading.Thread(target=pump, name="reader-thread", daemon=True)
        t.start()
        held = None
        try:
            while True:
                if held is not None:
                    eng.release_batch(held)
                    held = None
                nb = eng.next_batch()
                if nb is None:
                    break                                        # cancelled: the reader failed
                data, ends, status, err_off, off, first, last, held = nb
                if held == 0xFFFFFFFFFFFFFFFF:
                    held = None                                  # the end-of-input marker borrows nothing
                outcome = Outcome(status, status == 0, len(ends), 0, err_off, None)
                yield _Delimited(data, outcome, None, off, first, rel_ends=ends, at_eof=last and status == 0)
                if last:
                    break
        finally:
            eng.batch_cancel()
            t.join()
            if held is not None:
                eng.release_batch(held)
            eng.batch_end()
        if err:
            raise err[0]

    def ref_iter(self) -> RecordRefIter:
        return RecordRefIter(self._chunks())

    def each(self, func: Callable[[RefRecord], bool]) -> bool:
        """Apply func to every record; stop if it returns False.  Returns True at end of input,
        False if the closure stopped; raises FastqError -- after all earlier records were
        delivered -- on bad input."""
        it = self.ref_iter()
        try:
            while True:
                it.advance()
                rec = it.get()
                if rec is None:
                    return True
                if not func(rec):
                    return False
        finally:
            it.close()

    def record_sets(self) -> Iterator[RecordSet]:
        """The RecordSets the reference's RecordSetIter yields (src/lib.rs:364-425) -- same composition, same
        error behaviour -- rebuilt from the GPU's record index: a set = the records that are complete inside one
        fill of the reference's 68 KiB buffer.  With a reader that fills every read() the fills follow from the
        record boundaries alone (src/buffer.rs:30-48,74-100): the incomplete record is parked so that it ends on
        a 16-byte boundary and the next read adds n_free rounded down to 4096 (all of it below 4096).  The first
        set is empty (the buffer starts empty, src/lib.rs:381-391).  On bad input the records of the fill in
        which the error is DETECTED are dropped (src/lib.rs:375,399-410) and the error is raised; records of
        earlier fills have been yielded.  Every set owns a copy of its bytes, like the reference's."""
        fill_end = 0                       # stream offset up to which the reference's buffer has been filled
        pend: list = []                    # (bytes, start, e0..e3) of the records of the fill being built (stream offsets)
        first = True
        total_end = None                   # stream length, known at the end
        last = None
        for d in self._chunks():
            last = d
            n = len(d.starts)
            i = 0
            while i < n:
                s_abs, e_abs = d.base + int(d.starts[i]), d.base + int(d.ends[i, 3])
                if first:                  # the empty first set: the first fill happens inside the first next()
                    first = False
                    fill_end = BUFSIZE
                    yield RecordSet(memoryview(b""), np.empty(0, np.int64), np.empty((0, 4), np.int64))
                if e_abs < fill_end:       # complete inside the current fill
                    pend.append((bytes(d.buf[int(d.starts[i]):int(d.ends[i, 3]) + 1]), s_abs, d.ends[i] + d.base))
                    i += 1
                    continue
                # Incomplete (or EmptyBuffer at e_abs + 1 == ... ): the set goes out, the buffer is refilled
                yield _owned_set(pend)
                pend = []
                fill_end = _next_fill(fill_end, s_abs)
            if d.outcome.status != 0 or d.at_eof:
                break
        if last is None:
            return
        o = last.outcome
        if first:                          # no record at all: the reference still fills once (and that set is all
            fill_end = BUFSIZE             # there is if the input is empty: reader_at_end, src/lib.rs:386-388)
            yield RecordSet(memoryview(b""), np.empty(0, np.int64), np.empty((0, 4), np.int64))
            if o.status == 0 and len(last.buf) == 0:
                return
        stream_end = last.base + len(last.buf)        # (what the host has seen of the stream: enough for the checks below)
        if o.status == 0:
            # clean end: the pending set goes out when the buffer runs empty; one more (empty) set follows if that
            # refill still found bytes... it cannot at EOF: reader_at_end (src/lib.rs:386-388)
            yield _owned_set(pend)
            return
        # where the reference detects the error, relative to the fills
        e = o.err_offset
        while True:
            det = _detect_point(last, o.status, e)    # stream offset of the byte whose presence triggers the Err
            if o.status in (4, 5) or det is None or det >= fill_end:
                # Incomplete in this fill: the set goes out first (unless this very call raises too long / truncated)
                if o.status in (4, 5):
                    # too long: raised by the call that finds the buffer full (n_free() == 0 after parking it);
                    # truncated: by the call whose read returns 0.  Either way the records of that call are dropped
                    # -- but earlier fills that merely found the record incomplete went out.
                    nxt = _next_fill(fill_end, e)
                    if nxt == fill_end or fill_end >= stream_end:
                        break              # this call raises: pending records dropped
                    yield _owned_set(pend)
                    pend = []
                    fill_end = nxt
                    continue
                yield _owned_set(pend)
                pend = []
                fill_end = _next_fill(fill_end, e)
                continue
            break                          # detected inside the current fill: its records are dropped
        last.raise_for_status()

    def parallel_each(self, n_threads: int, func: Callable[[Iterator[RecordSet]], object]) -> list:
        """n_threads workers, each fed RecordSets round-robin over a bounded channel of 10
        (src/lib.rs:509-566).  Returns the workers' results in worker order; raises FastqError on
        bad input (after joining the workers).  A worker that returns before its iterator is exhausted
        hangs up its channel: the producer's next send to it fails and the producer stops, as in the
        reference (src/lib.rs:540-542, doc-test "Early return stops the parser" :484)."""
        chans = [_SyncChannel(10) for _ in range(n_threads)]
        results: list = [None] * n_threads
        errors: list = [None] * n_threads

        def worker(i):
            try:
                results[i] = func(chans[i].receive())
            except BaseException as e:  # worker panic -> re-raised on join (src/lib.rs:558)
                errors[i] = e
            finally:
                chans[i].hang_up()      # rx dropped: pending and future sends fail

        threads = [threading.Thread(target=worker, args=(i,), name=f"worker-{i}") for i in range(n_threads)]
        for t in threads:
            t.start()
        err = None
        try:
            if n_threads:
                for k, s in enumerate(self.record_sets()):
                    if not chans[k % n_threads].send(s):
                        break           # the worker quit: stop parsing (src/lib.rs:540-542)
        except FastqError as e:
            err = e
        except BaseException:
            for c in chans:
                c.close()
            raise
        for c in chans:
            c.close()                   # drop(senders): the workers' iterators end (src/lib.rs:551)
        for t in threads:
            t.join()
        for e in errors:
            if e is not None:
                raise e
        if err is not None:
            raise err
        return results

    # ---- fast paths: nothing but counts / histograms ever reaches the host -------------------
    def _stream(self, hist: bool):
        """Feed the reader through the pinned ring (the thread_reader protocol, src/thread_reader.rs:
        40-50 and 126-147): a reader thread fills pinned slots straight from `reader.readinto` --
        acquire = empty_recv.recv(), submit = full_send.send() -- while the copy stream and the
        kernels drain the slots behind it.  In-memory inputs skip the thread and go through
        fqb_parse_host."""
        r = self._reader
        if isinstance(r, (bytes, bytearray, memoryview, np.ndarray)):
            data = _read_all(r)
            outcome, st, _ = self._engine.parse_host(data, hist=hist, want_stats=hist)
            return outcome, st
        eng = self._engine
        eng.stream_begin(hist=hist)
        err: list = []

        def pump():
            try:
                while True:
                    slot = eng.stream_acquire()                  # blocks until a pinned slot is free
                    mv = memoryview(slot).cast("B")
                    if hasattr(r, "readinto"):
                        n = r.readinto(mv)                       # one read per slot, may be short
                    else:
                        b = r.read(len(mv))
                        n = len(b)
                        mv[:n] = b
                    eng.stream_submit(n or 0)
                    if not n:
                        return
            except BaseException as e:                           # reader errors travel to the caller
                err.append(e)

        t = threading.Thread(target=pump, name="reader-thread")
        t.start()
        t.join()
        if err:
            try:
                eng.stream_finish(want_stats=False)
            finally:
                raise err[0]
        return eng.stream_finish(want_stats=hist)

    def count(self) -> int:
        """examples/fastq-count.rs: number of records (raises on bad input)."""
        outcome, _ = self._stream(hist=False)
        outcome.raise_for_status()
        return outcome.n_records

    def stats(self) -> tuple[Outcome, Stats]:
        """The stats closure (per-position base / quality histograms over seq()/qual()) run on the
        GPU; Outcome.status tells whether (and where) each() would have returned Err."""
        return self._stream(hist=True)


    def filter_to(self, writer, keep: str = "dnan") -> int:
        """Write the records whose seq() passes validate_dna ("dna") / validate_dnan ("dnan") -- or
        every record ("all") -- to `writer`, verbatim (Record::write, src/records.rs:93-96) and in
        order; predicate and compaction run on the GPU.  Returns the number of records written;
        raises FastqError after the records in front of a bad one have been written (each()'s order).
        (The whole input is staged in HBM at once: inputs larger than the GPU are fed shard by shard
        through Engine.parse_device / Engine.filter_device.)"""
        import torch
        from ._lib import KEEP_ALL, KEEP_DNA, KEEP_DNAN
        mode = {"all": KEEP_ALL, "dna": KEEP_DNA, "dnan": KEEP_DNAN}[keep]
        eng = self._engine
        data = _read_all(self._reader)
        dev = f"cuda:{eng.device}"
        d = torch.zeros(data.size + 64, dtype=torch.uint8, device=dev)
        if data.size:
            d[:data.size] = torch.from_numpy(data if data.flags.writeable else data.copy())
        idx = torch.empty(data.size + 8, dtype=torch.int32, device=dev)
        eng.parse_device(d, n_own=data.size, n_avail=data.size, hist=False, index=idx)
        outcome, _ = eng.fetch(want_stats=False)
        out = torch.empty(max(data.size, 1), dtype=torch.uint8, device=dev)
        eng.filter_device(d, idx, outcome.n_records, mode, out)
        n_kept, n_bytes = eng.fetch_filter()
        writer.write(out[:n_bytes].cpu().numpy().tobytes())
        outcome.raise_for_status()
        return n_kept


def each_zipped(parser1: Parser, parser2: Parser, callback) -> tuple[bool, bool]:
    """src/lib.rs:577-609, lock-step over two delimited streams."""
    it1, it2 = parser1.ref_iter(), parser2.ref_iter()
    finished = (False, False)
    try:
        it1.advance()
        it2.advance()
        while True:
            v1 = None if finished[0] else it1.get()
            v2 = None if finished[1] else it2.get()
            finished = (v1 is None, v2 is None)
            adv = callback(v1, v2)
            if tuple(adv) == (False, False) or finished == (True, True):
                return finished
            if adv[0] and not finished[0]:
                it1.advance()
            if adv[1] and not finished[1]:
                it2.advance()
    finally:
        it1.close()
        it2.close()


def parse_path(path, func, max_len: int = 150):
    """src/lib.rs:167-196 for uncompressed input: open `path` (None / '-' = stdin) and hand a
    Parser to func.  Decompression (niffler) is outside the accelerated path."""
    import sys
    if path is None or path == "-":
        return func(Parser(sys.stdin.buffer, max_len=max_len))
    with open(path, "rb") as f:
        magic = f.read(4)
        f.seek(0)
        if magic[:2] == b"\x1f\x8b" or magic[:3] == b"BZh" or magic[:4] in (b"\xfd7zX", b"\x04\"M\x18"):
            raise FastqError(6, 0, 0)
        return func(Parser(f, max_len=max_len))
