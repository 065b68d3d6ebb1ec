"""Pins the CPU oracle (oracle/fastq_oracle.c) to the reference's own unit tests
(tests/golden/reference_unit_tests.json, transcribed from src/lib.rs:611-811 and the
doc-test at src/lib.rs:474-508), to the edge-case list of SURVEY.md 8(a), and fuzzes it
against the independent pure-Python model in tests/pymodel.py."""
import base64
import json
import os

import numpy as np
import pytest
from hypothesis import given, settings, strategies as st

import pymodel
from oracle import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
with open(os.path.join(HERE, "golden", "reference_unit_tests.json")) as f:
    GOLD = json.load(f)


def d64(s):
    return base64.b64decode(s)


def vec_input(v) -> bytes:
    if "input" in v:
        return d64(v["input"])
    g = v["input_gen"]
    return d64(g["prefix"]) + d64(g["unit"]) * g["times"] + d64(g["suffix"])


def check_records(recs, expect):
    assert len(recs) == len(expect)
    for r, e in zip(recs, expect):
        assert r.head == d64(e["head"])
        assert r.seq == d64(e["seq"])
        assert r.qual == d64(e["qual"])
        if "write" in e:
            assert r.raw == d64(e["write"])  # RefRecord::write == raw bytes, records.rs:93-96
            # OwnedRecord::write (records.rs:113-128) reassembles the same bytes for LF input
            owned = b"@" + r.head + b"\n" + r.seq + b"\n" + r.sep + b"\n" + r.qual + b"\n"
            assert owned == d64(e["write"])


@pytest.mark.parametrize("v", GOLD["vectors"], ids=[v["name"] for v in GOLD["vectors"]])
def test_reference_unit_test(v):
    data = vec_input(v)
    exp = v["expect"]
    if v["api"] == "each":
        res, recs = oracle.each(data)
        assert (res.status == oracle.OK) == exp["ok"]
        if exp["ok"]:
            assert res.finished
        if "records" in exp:
            check_records(recs, exp["records"])
        if "expect_kind" in v:
            assert res.error == v["expect_kind"]
    elif v["api"] == "record_sets":
        status, sets = oracle.record_sets(data)
        assert (status == oracle.OK) == exp["ok"]
        flat = [r for s in sets for r in s]
        if "count" in exp:
            assert len(flat) == exp["count"]
        if "records" in exp:
            check_records(flat, exp["records"])
        if "expect_kind" in v:
            assert oracle.ERR_NAMES[status] == v["expect_kind"]
    elif v["api"] == "parallel_each":
        rc, n = oracle.parallel_each_count(data, v["n_threads"])
        assert (rc == oracle.OK) == exp["ok"]
        assert n == exp["count"]
        rc2, st_, _sets = oracle.parallel_each_stats(data, 16, v["n_threads"])
        assert rc2 == rc and st_.n_records == exp["count"]
        if "records" in exp:
            _res, recs = oracle.each(data)
            check_records(recs, exp["records"])
    else:
        raise AssertionError(v["api"])


# ---- SURVEY.md 8(a) edge-case list (derived, not asserted by the reference's tests) -------
EDGE = [
    (b"", "ok", 0),
    (b"\n", "header", 0),
    (b"@hi\nNN\n+\n++\n\n", "header", 1),                 # trailing blank line
    (b"@id\n\n+\n\n", "ok", 1),                            # empty sequence
    (b"@hi\nNN\r\n+\n++\n", "length", 0),                  # mixed endings, raw lengths differ
    (b"@hi\nNN\r\n+\n+++\n", "ok", 1),                     # raw lengths equal, views differ
    (b"@id\nAC\n+\n@+\n@i2\nGT\n+\n+@\n", "ok", 2),        # '@'/'+' leading quality lines
    (b"@hi\nNN\n+\n++\n@hi\nNN\nX", "sep", 1),             # partial final record, bad sep
    (b"@hi\nNN\n+\n++\nhi", "header", 1),                  # partial final record, bad header
    (b"@hi\nNN\n+\n++\n@hi\nNN\n", "truncated", 1),
    (b"@hi\nNN\n+\n++\n@hi\nNN\n+", "truncated", 1),
    (b"@hi\nNN\nxx\n++\n", "sep", 0),
    (b"hi\nNN\n+\n++\n", "header", 0),
]


@pytest.mark.parametrize("data,kind,n", EDGE)
def test_edge_cases(data, kind, n):
    for bufsize, max_read in ((oracle.BUFSIZE, 0), (64, 1), (128, 7), (4096, 13)):
        res, recs = oracle.each(data, bufsize=bufsize, max_read=max_read)
        assert res.error == kind, (bufsize, max_read)
        assert len(recs) == n and res.n_records == n
    st_, recs_m = pymodel.each(data)
    assert oracle.ERR_NAMES[st_] == kind and len(recs_m) == n


def test_crlf_views_and_mixed_lengths():
    res, recs = oracle.each(b"@hi\nNN\r\n+\n+++\n")
    assert res.status == oracle.OK
    assert recs[0].seq == b"NN" and recs[0].qual == b"+++"
    res, recs = oracle.each(b"@a\n\r\n+\n\r\n")
    assert recs[0].seq == b"" and recs[0].qual == b""


def test_too_long_thresholds():
    B = oracle.BUFSIZE
    # at stream start a record of exactly BUFSIZE bytes is accepted (test `bufflen`)
    def rec_of(total):
        return b"@" + b"a" * (total - 8) + b"\nA\n+\nB\n"
    assert oracle.each(rec_of(B))[0].status == oracle.OK
    assert oracle.each(rec_of(B + 1))[0].error == "too_long"
    # always accepted at <= BUFSIZE-15 whatever precedes it
    for lead in (0, 1, 5, 16, 321):
        pre = b"@x\nA\n+\nB\n" * lead
        r = oracle.each(pre + rec_of(B - 15))[0]
        assert r.status == oracle.OK and r.n_records == lead + 1
        assert oracle.each(pre + rec_of(B + 1))[0].error == "too_long"


def test_validate_dna():
    _res, recs = oracle.each(b"@a\nACGT\n+\n!!!!\n@b\nACGN\n+\n!!!!\n@c\nacgt\n+\n!!!!\n")
    assert [(r.valid_dna, r.valid_dnan) for r in recs] == [(True, True), (False, True),
                                                           (False, False)]


def test_each_filter_matches_the_python_model():
    """fqo_each_filter = each() + `if rec.validate_dna(n)() { rec.write(w) }` (src/records.rs:19-33, 93-96),
    checked against the independent pure-Python model of each() with the predicate spelled out."""
    import numpy as np
    rng = np.random.default_rng(3)
    recs = []
    for i in range(400):
        L = int(rng.integers(0, 40))
        alpha = b"ACGT" if i % 3 else (b"ACGTN" if i % 2 else b"ACGTNacgtX.")
        seq = bytes(rng.choice(np.frombuffer(alpha, dtype=np.uint8), L))
        e = b"\r\n" if i % 7 == 0 else b"\n"
        recs.append(b"@r%d" % i + e + seq + e + b"+" + e + b"I" * L + e)
    for data in (b"".join(recs), b"".join(recs) + b"@x\nAC", b"", b"@a\n\n+\n\n", b"".join(recs[:50]) + b"junk\n"):
        status, model = pymodel.each(data)
        for mode, alphabet in ((0, None), (1, b"ACTG"), (2, b"ACTGN")):
            keep = [r for r in model if alphabet is None or all(c in alphabet for c in r[1])]
            res, n_kept, out = oracle.each_filter(data, mode)
            assert res.status == status and res.n_records == len(model)
            assert n_kept == len(keep)
            assert out == b"".join(r[3] for r in keep)


def test_each_stop_early():
    data = b"@a\nA\n+\n!\n" * 5
    seen = []
    res, _ = oracle.each(data, callback=lambda r: (seen.append(r.offset), len(seen) < 3)[1])
    assert res.status == oracle.OK and not res.finished and len(seen) == 3


def test_first_record_set_is_empty():
    status, sets = oracle.record_sets(b"@hi\nNN\n+\n++\n")
    assert status == oracle.OK and len(sets[0]) == 0 and sum(map(len, sets)) == 1


# ---- synthetic generators -----------------------------------------------------------------
def test_synth_fixed_format():
    a = oracle.synth_fixed_records(31152)
    assert a.size == 9_999_792  # config 1 file
    first = bytes(a[:321])
    assert first.startswith(b"@FQ0000000000000\n") and first[167:170] == b"\n+\n" and first[320] == 10
    res, st_ = oracle.each_stats(a, 150)
    assert res.status == oracle.OK and st_.n_records == 31152 and st_.n_bases == 31152 * 150
    assert st_.len_hist[150] == 31152
    assert st_.base_hist.sum() == 31152 * 150 and st_.qual_hist.sum() == 31152 * 150
    assert st_.base_hist[:, 5].sum() == 0 and st_.base_hist[:, 4].sum() > 0
    q = np.nonzero(st_.qual_hist.sum(axis=0))[0]
    assert q.min() == 35 and q.max() == 74  # '#'..'J'
    # windowed generation is consistent with whole-stream generation
    w = oracle.synth_fixed(5000, 150, byte_off=123457)
    assert bytes(w) == bytes(a[123457:123457 + 5000])


def test_synth_var_format():
    a = oracle.synth_var(2000)
    res, st_ = oracle.each_stats(a, 300)
    assert res.status == oracle.OK and st_.n_records == 2000
    lens = np.nonzero(st_.len_hist)[0]
    assert lens.min() >= 50 and lens.max() <= 300 and st_.clip_seq == 0
    b = oracle.synth_var(500, first=700)
    off = sum(17 + 2 * (oracle.synth_var_len(i) + 1) + 2 for i in range(700))
    assert bytes(b) == bytes(a[off:off + b.size])


def test_parallel_each_matches_each():
    a = oracle.synth_var(5000)
    res, s1 = oracle.each_stats(a, 300)
    for n in (1, 3):
        rc, s2, sets = oracle.parallel_each_stats(a, 300, n)
        assert rc == oracle.OK and s2.n_records == s1.n_records == res.n_records
        assert np.array_equal(s1.base_hist, s2.base_hist)
        assert np.array_equal(s1.qual_hist, s2.qual_hist)
        assert np.array_equal(s1.len_hist, s2.len_hist)
        assert sets.sum() > 1


# ---- fuzz: C oracle vs independent Python model; chunking independence ---------------------
line = st.binary(min_size=0, max_size=12).map(lambda b: b.replace(b"\n", b"A"))


@st.composite
def fastq_like(draw):
    n = draw(st.integers(0, 6))
    out = b""
    for _ in range(n):
        L = draw(st.integers(0, 9))
        seq = bytes(draw(st.lists(st.sampled_from(b"ACGTNacgt@+\r"), min_size=L, max_size=L)))
        qual = bytes(draw(st.lists(st.integers(33, 126), min_size=L, max_size=L)))
        eol = draw(st.sampled_from([b"\n", b"\r\n"]))
        out += b"@" + draw(line) + eol + seq + eol + b"+" + draw(line) + eol + qual + eol
    # mutate
    for _ in range(draw(st.integers(0, 3))):
        if not out:
            break
        i = draw(st.integers(0, len(out) - 1))
        op = draw(st.integers(0, 2))
        if op == 0:
            out = out[:i] + out[i + 1:]
        elif op == 1:
            out = out[:i] + bytes([draw(st.sampled_from(b"\n@+A\r"))]) + out[i:]
        else:
            out = out[:i]
    return out


@settings(max_examples=400, deadline=None)
@given(fastq_like(), st.sampled_from([(64, 1), (64, 0), (128, 7), (4096, 13), (68 * 1024, 0)]))
def test_fuzz_oracle_vs_pymodel(data, cfg):
    bufsize, max_read = cfg
    res, recs = oracle.each(data, bufsize=bufsize, max_read=max_read)
    st_m, recs_m = pymodel.each(data, bufsize=bufsize, max_read=max_read)
    assert res.status == st_m
    assert [(r.head, r.seq, r.qual, r.raw, r.offset) for r in recs] == recs_m
    # stats closure agrees with the python model as well
    _r, s = oracle.each_stats(data, 8, bufsize=bufsize, max_read=max_read)
    m = pymodel.stats(recs_m, 8)
    assert (s.n_records, s.n_bases, s.clip_seq, s.clip_qual) == (
        m["n_records"], m["n_bases"], m["clip_seq"], m["clip_qual"])
    assert np.array_equal(s.base_hist, m["base"]) and np.array_equal(s.qual_hist, m["qual"])
    assert np.array_equal(s.len_hist, m["lens"])
    # index form
    _r, idx = oracle.each_index(data, bufsize=bufsize, max_read=max_read)
    assert [int(x) for x in idx[:, 0]] == [r.offset for r in recs]


def _rec_total(n: int) -> bytes:
    return b"@" + b"a" * (n - 8) + b"\nA\n+\nB\n"


@pytest.mark.parametrize("first", [[], [8], [15], [16], [45], [1000, 13], [69000], [40000, 40000, 11]],
                         ids=lambda f: "+".join(map(str, f)) or "start")
def test_too_long_band_closed_form(first):
    """The closed form the CUDA path implements for "Fastq record is too long" (include/fastq_b200.h:
    a record at stream offset p fits iff (p mod 16) + length <= BUFSIZE; an incomplete one is too long iff the
    stream holds BUFSIZE - p mod 16 bytes from p on, else truncated) against the oracle's restatement of
    Buffer::clean / read_into / RecordRefIter::advance (src/buffer.rs:51-100, src/lib.rs:255-303), for a reader
    that fills every read -- and the pure-Python model agrees on a subset."""
    BUF = 68 * 1024
    pre = b"".join(_rec_total(n) for n in first)
    p = len(pre)
    for L in range(69605, 69645):
        data = pre + _rec_total(L) + _rec_total(20)
        res, _ = oracle.each_index(data)
        fits = (p & 15) + L <= BUF
        assert (res.status, res.n_records) == ((0, len(first) + 2) if fits else (4, len(first))), (first, L)
        if L in (69616, 69617, 69625, 69632, 69633):
            st_m, recs_m = pymodel.each(data)
            assert (st_m, len(recs_m)) == (res.status, res.n_records)
    for avail in range(69605, 69645):
        res, _ = oracle.each_index(pre + b"@" + b"a" * (avail - 1))
        assert (res.status, res.n_records) == (4 if avail >= BUF - (p & 15) else 5, len(first)), (first, avail)
