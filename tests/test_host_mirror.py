"""Host mirror of the crate's drivers (fastq_rs_b200/parser.py) on CPU: the delimiting engine is replaced by an
oracle-backed stand-in (tests/fake_engine.py: FakeHostEngine), so these tests cover the HOST logic only --
refills with carry-over, RecordRefIter, record_sets, parallel_each's channels, each_zipped -- against the oracle
and against small pure-Python models of the reference's control flow.  The GPU suite runs the same drivers over
the real engine (tests/test_gpu_parity.py)."""
import io
import threading
import time

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

from fake_engine import FakeHostEngine
from fastq_rs_b200 import parser as fqp
from fastq_rs_b200.engine import FastqError
from oracle import oracle


def _rec(i, L):
    rng = np.random.default_rng(i)
    seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), L))
    qual = bytes(rng.integers(33, 75, L, dtype=np.uint8))
    return b"@r%d\n" % i + seq + b"\n+\n" + qual + b"\n"


def _parser(data, chunk=None, **kw):
    return fqp.Parser(io.BytesIO(data), engine=FakeHostEngine(), chunk_bytes=chunk or fqp.CHUNK_BYTES, **kw)


@pytest.mark.parametrize("chunk", [1, 100, 4097, 1 << 20])
def test_each_over_refills_matches_the_oracle(chunk):
    recs = [_rec(i, L) for i, L in enumerate([150, 0, 3, 150, 5000, 151, 1, 150] * (1 if chunk == 1 else 3))]
    good = b"".join(recs)
    for data in (good, good[:-1], good + b"\n", good[:len(good) // 2] + b"X" + good[len(good) // 2:], b"", b"\n"):
        ores, oidx = oracle.each_index(data)
        seen, err = [], None
        try:
            assert _parser(data, chunk).each(lambda r: seen.append((bytes(r.data), r.head(), r.seq(), r.qual())) or True)
        except FastqError as e:
            err = e
        assert len(seen) == ores.n_records and (err.status if err else 0) == ores.status
        if err:
            assert (err.offset, err.n_delivered) == (ores.err_offset, ores.n_records)
        for (raw, head, seq, qual), (s, e0, e1, e2, e3) in zip(seen, oidx.astype(int)):
            assert raw == data[s:e3 + 1] and head == data[s + 1:e0] and seq == data[e0 + 1:e1] and qual == data[e2 + 1:e3]


def test_each_stops_when_the_closure_returns_false():
    data = b"".join(_rec(i, 20) for i in range(100))
    n = []
    assert _parser(data, 500).each(lambda r: n.append(1) or len(n) < 7) is False
    assert len(n) == 7


def test_parallel_each_early_return_stops_the_parser():
    """src/lib.rs:484 ('Early return stops the parser') with far more than 10 * n_threads record sets pending:
    a worker that returns at once hangs up its channel, the producer's next send to it fails, the producer
    stops (src/lib.rs:540-542) -- parallel_each returns instead of blocking on the full queue of a dead worker."""
    data = b"".join(_rec(i, 150) for i in range(40000))          # ~190 record sets of 68 KiB
    assert len(data) > 100 * fqp.BUFSIZE
    box = {}

    def run():
        box["out"] = _parser(data).parallel_each(2, lambda sets: next(iter(sets), None) is not None)
    t = threading.Thread(target=run, daemon=True)
    t0 = time.time()
    t.start()
    t.join(timeout=60)
    assert not t.is_alive(), "parallel_each deadlocked on a worker that returned early"
    assert box["out"] == [True, True] and time.time() - t0 < 60


def test_parallel_each_one_worker_quits_the_other_keeps_its_sets():
    data = b"".join(_rec(i, 150) for i in range(20000))
    def work(sets):
        n = 0
        for s in sets:
            n += s.len()
            if threading.current_thread().name == "worker-0":
                return -1                                           # quits after its first set
        return n
    out = _parser(data).parallel_each(2, work)
    assert out[0] == -1 and 0 <= out[1] < 20000


def test_parallel_each_counts_everything_and_reports_errors():
    data = b"".join(_rec(i, 150) for i in range(5000))
    out = _parser(data, 300000).parallel_each(3, lambda sets: sum(s.len() for s in sets))
    assert sum(out) == 5000 and len(out) == 3
    with pytest.raises(FastqError) as ei:
        _parser(data + b"@x\nAC\n+\n!\n", 300000).parallel_each(3, lambda sets: sum(s.len() for s in sets))
    assert ei.value.status == oracle.E_LENGTH
    with pytest.raises(ZeroDivisionError):                           # worker panic -> re-raised on join
        _parser(data).parallel_each(2, lambda sets: 1 // 0)


# ---- each_zipped against a transliteration-free model of src/lib.rs:577-609 ---------------------
def _zip_model(n1, n2, flags):
    """What the reference does with two streams of n1 / n2 records and a callback answering flags[k] at its
    k-th call: the list of (index1 | None, index2 | None) pairs the callback sees, and the result."""
    i1 = i2 = 0                      # records consumed so far = index of the current one
    fin = (False, False)
    calls = []
    k = 0
    while True:
        v1 = None if fin[0] or i1 >= n1 else i1
        v2 = None if fin[1] or i2 >= n2 else i2
        fin = (v1 is None, v2 is None)
        calls.append((v1, v2))
        adv = flags[k % len(flags)]
        k += 1
        if adv == (False, False) or fin == (True, True):
            return calls, fin
        if adv[0] and not fin[0]:
            i1 += 1
        if adv[1] and not fin[1]:
            i2 += 1


@settings(max_examples=60, deadline=None)
@given(st.integers(0, 40), st.integers(0, 40),
       st.lists(st.tuples(st.booleans(), st.booleans()), min_size=1, max_size=12),
       st.sampled_from([64, 1000, 1 << 20]), st.sampled_from([77, 5000, 1 << 20]))
def test_each_zipped_against_the_model(n1, n2, flags, c1, c2):
    if all(f == (False, True) for f in flags) or all(f == (True, False) for f in flags):
        flags = flags + [(True, True)]            # (a callback that never advances one side loops for ever, as in the reference, unless the other side ends it)
    a = b"".join(_rec(i, 5 + i % 7) for i in range(n1))
    b = b"".join(_rec(1000 + i, 3 + i % 5) for i in range(n2))
    calls, k = [], [0]

    def cb(r1, r2):
        calls.append((None if r1 is None else r1.head(), None if r2 is None else r2.head()))
        f = flags[k[0] % len(flags)]
        k[0] += 1
        return f
    want_calls, want_fin = _zip_model(n1, n2, flags)
    if len(want_calls) > 2000:
        return
    fin = fqp.each_zipped(_parser(a, c1), _parser(b, c2), cb)
    assert fin == want_fin
    assert calls == [(None if i is None else b"r%d" % i, None if j is None else b"r%d" % (1000 + j)) for i, j in want_calls]


def test_each_zipped_propagates_errors_of_either_stream():
    good = b"".join(_rec(i, 10) for i in range(10))
    bad = good[:60] + b"X" + good[60:]
    for a, b in ((good, bad), (bad, good)):
        with pytest.raises(FastqError):
            fqp.each_zipped(_parser(a, 50), _parser(b, 50), lambda r1, r2: (True, True))


# ---- record_sets: same sets, same dropped records, same error as the reference's RecordSetIter --------------
def _sets_of(parser):
    out, err = [], 0
    try:
        for s in parser.record_sets():
            out.append([bytes(r.data) for r in s.iter()])
    except FastqError as e:
        err = e.status
    return err, out


def _long(i, n):
    """a record of exactly n bytes"""
    return b"@" + b"h" * (n - 8) + b"\nA\n+\nB\n"


RS_CASES = {
    "empty": b"",
    "one": _rec(0, 150),
    "many_fixed": b"".join(_rec(i, 150) for i in range(1500)),
    "many_var": b"".join(_rec(i, 10 + (i * 37) % 400) for i in range(1200)),
    "exactly_bufsize": _long(0, 68 * 1024),
    "bufsize_then_more": _long(0, 68 * 1024) + b"".join(_rec(i, 99) for i in range(300)),
    "long_records": b"".join(_long(i, n) for i, n in enumerate([30000, 40000, 69000, 8, 69616, 50, 69600])),
    "too_long_mid": b"".join(_rec(i, 150) for i in range(400)) + _long(0, 70000) + _rec(1, 5),
    "too_long_band": _rec(0, 7) + _long(0, 69625) + _rec(1, 5),
    "truncated": b"".join(_rec(i, 150) for i in range(500))[:-1],
    "truncated_tail_record": b"".join(_rec(i, 150) for i in range(500)) + b"@tail\nACGT\n",
    "bad_header_mid": b"".join(_rec(i, 150) for i in range(700)) + b"X" + b"".join(_rec(i, 150) for i in range(50)),
    "bad_sep": b"".join(_rec(i, 150) for i in range(650)) + b"@x\nACGT\n-\nIIII\n" + _rec(9, 9),
    "bad_len": b"".join(_rec(i, 150) for i in range(333)) + b"@x\nACGT\n+\nIII\n" + _rec(9, 9),
    "blank_line_end": b"".join(_rec(i, 150) for i in range(100)) + b"\n",
}


@pytest.mark.parametrize("name", sorted(RS_CASES))
@pytest.mark.parametrize("chunk", [None, 100000, 4097])
def test_record_sets_match_the_reference_fills(name, chunk):
    data = RS_CASES[name]
    ostatus, osets = oracle.record_sets(data)
    err, sets = _sets_of(_parser(data, chunk))
    assert err == ostatus, (name, err, ostatus)
    assert [len(x) for x in sets] == [len(x) for x in osets], name
    assert [[r for r in x] for x in sets] == [[r.raw for r in x] for x in osets]


@settings(max_examples=40, deadline=None)
@given(st.lists(st.integers(0, 1200), min_size=0, max_size=300), st.integers(0, 3), st.sampled_from([None, 50000]))
def test_record_sets_fuzz_against_the_oracle(lens, damage, chunk):
    recs = [_rec(i, L) for i, L in enumerate(lens)]
    data = b"".join(recs)
    if damage == 1 and data:
        data = data[:len(data) * 2 // 3]
    elif damage == 2 and len(recs) > 2:
        k = len(b"".join(recs[:len(recs) // 2]))
        data = data[:k] + b"#" + data[k + 1:]
    elif damage == 3 and len(recs) > 2:
        k = len(b"".join(recs[:len(recs) // 2])) + recs[len(recs) // 2].index(b"\n+\n") + 1
        data = data[:k] + b"-" + data[k + 1:]
    ostatus, osets = oracle.record_sets(data)
    err, sets = _sets_of(_parser(data, chunk))
    assert err == ostatus
    assert [[r for r in x] for x in sets] == [[r.raw for r in x] for x in osets]
