"""GPU parity tests: the sm_100a path (through the C ABI) against the CPU oracle.

Bar: bit-exact for record counts, record offsets, error kind + offset, and every histogram bin.
Covers the reference's own unit tests (tests/golden/reference_unit_tests.json, src/lib.rs:611-811),
the edge-case list of SURVEY.md 8(a), synthetic inputs A/B of 8(d), shard cuts at every byte,
the streaming ring with chunk boundaries, and hypothesis-mutated inputs.
"""
import base64
import io
import json
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

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


@pytest.fixture(scope="module")
def torch():
    import torch
    return torch


@pytest.fixture(scope="module")
def oracle():
    from oracle import oracle
    return oracle


@pytest.fixture(scope="module")
def fq():
    import fastq_rs_b200 as fq
    return fq


@pytest.fixture(scope="module")
def eng(fq):
    e = fq.Engine(max_len=150, slot_bytes=1 << 20)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng300(fq):
    e = fq.Engine(max_len=300, slot_bytes=1 << 20)
    yield e
    e.close()


@pytest.fixture(scope="module")
def eng16(fq):
    e = fq.Engine(max_len=16, slot_bytes=1 << 20)
    yield e
    e.close()


def to_dev(torch, data: bytes, pad: int = 64):
    a = np.frombuffer(data, dtype=np.uint8)
    t = torch.zeros(len(data) + pad, dtype=torch.uint8, device="cuda")
    if len(data):
        t[:len(data)] = torch.from_numpy(a.copy())
    return t


def assert_stats_equal(st, ost):
    assert st.n_records == ost.n_records
    assert st.n_bases == ost.n_bases
    assert st.clip_seq == ost.clip_seq
    assert st.clip_qual == ost.clip_qual
    np.testing.assert_array_equal(st.len_hist, ost.len_hist)
    np.testing.assert_array_equal(st.base_hist, ost.base_hist)
    np.testing.assert_array_equal(st.qual_hist, ost.qual_hist)


def check_device_vs_oracle(torch, oracle, engine, data: bytes, check_index=True):
    """parse_device over the whole buffer == oracle.each on the same bytes."""
    P = engine.max_len
    t = to_dev(torch, data)
    idx = torch.zeros(len(data) + 8, dtype=torch.int32, device="cuda")
    engine.parse_device(t, n_own=len(data), n_avail=len(data), hist=True, index=idx)
    out, st = engine.fetch()
    ores, ost = oracle.each_stats(data, P)
    assert out.status == ores.status, (out, ores)
    assert out.n_records == ores.n_records
    if ores.status != 0:
        assert out.err_offset == ores.err_offset
        assert not out.finished
    else:
        assert out.finished
    assert out.n_lines == data.count(b"\n")
    assert_stats_equal(st, ost)
    if check_index:
        _, oidx = oracle.each_index(data)
        n = out.n_records
        got = idx[:4 * n].cpu().numpy().view(np.uint32).astype(np.uint64).reshape(n, 4)
        np.testing.assert_array_equal(got, oidx[:, 1:5])
        if n:
            starts = np.concatenate([[0], got[:-1, 3] + 1]).astype(np.uint64)
            np.testing.assert_array_equal(starts, oidx[:, 0])
    # the host path (pinned ring, chunked) must agree too
    hout, hst, hidx = engine.parse_host(data, want_index=True)
    assert (hout.status, hout.n_records) == (ores.status, ores.n_records)
    if ores.status != 0:
        assert hout.err_offset == ores.err_offset
    assert_stats_equal(hst, ost)
    return out, st


# --------------------------------------------------------------------------------------------
# the reference's own unit tests
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("v", GOLD["vectors"], ids=[v["name"] for v in GOLD["vectors"]])
def test_reference_unit_test(v, fq, torch, oracle, eng):
    data = vec_input(v)
    exp = v["expect"]
    recs = []
    err = None
    try:
        if v["api"] == "each":
            finished = fq.Parser(data, engine=eng).each(lambda r: recs.append(r.to_owned_record()) or True)
            assert finished
        elif v["api"] == "record_sets":
            for s in fq.Parser(data, engine=eng).record_sets():
                recs.extend(r.to_owned_record() for r in s)
        elif v["api"] == "parallel_each":
            counts = fq.Parser(data, engine=eng).parallel_each(
                v["n_threads"], lambda sets: sum(s.len() for s in sets))
            assert sum(counts) == exp["count"]
    except fq.FastqError as e:
        err = e
    assert (err is None) == exp["ok"]
    if err is not None:
        assert err.kind == "InvalidData"
        if "expect_kind" in v:
            assert oracle.ERR_NAMES[err.status] == v["expect_kind"]
    if "records" in exp and v["api"] != "parallel_each":
        assert len(recs) == len(exp["records"])
        for r, e in zip(recs, exp["records"]):
            assert r.head() == d64(e["head"]) and r.seq() == d64(e["seq"]) and r.qual() == d64(e["qual"])
            if "write" in e:
                w = io.BytesIO()
                assert r.write(w) == len(d64(e["write"]))
                assert w.getvalue() == d64(e["write"])
    if "count" in exp and v["api"] == "record_sets" and exp["ok"]:
        assert len(recs) == exp["count"]
    check_device_vs_oracle(torch, oracle, eng, data)


def test_refrecord_write_verbatim(fq, eng):
    data = b"@hi\r\nNN\r\n+hi\r\n++\r\n"
    out = []
    fq.Parser(data, engine=eng).each(lambda r: out.append((r.head(), r.seq(), r.qual(), bytes(r.data))) or True)
    assert out == [(b"hi", b"NN", b"++", data)]


def test_each_stops_when_closure_returns_false(fq, eng, oracle):
    data = oracle.synth_fixed_records(10).tobytes()
    seen = []
    assert fq.Parser(data, engine=eng).each(lambda r: seen.append(1) or len(seen) < 3) is False
    assert len(seen) == 3


# --------------------------------------------------------------------------------------------
# edge cases
# --------------------------------------------------------------------------------------------
EDGE = {
    "empty": b"",
    "lone_newline": b"\n",
    "trailing_blank_line": b"@a\nAC\n+\n!!\n\n",
    "empty_seq": b"@id\n\n+\n\n",
    "empty_seq_many": b"@\n\n+\n\n" * 50,
    "crlf": b"@hi\r\nNN\r\n+\r\n++\r\n@ho\r\nACGT\r\n+\r\n!!!!\r\n",
    "mixed_mismatch": b"@hi\nNN\r\n+\n++\n",
    "mixed_accepted": b"@hi\nNN\r\n+\n+++\n",
    "lone_cr_lines": b"@hi\n\r\n+\n\r\n",
    "qual_starts_with_at": b"@id\nAC\n+\n@+\n@id2\nGT\n+\n+@\n",
    "partial_sep_error": b"@a\nAC\n+\n!!\n@hi\nNN\nX",
    "partial_header_error": b"@a\nAC\n+\n!!\nhi",
    "partial_truncated": b"@a\nAC\n+\n!!\n@hi\nNN\n",
    "partial_truncated2": b"@a\nAC\n+\n!!\n@hi\nNN\n+",
    "no_final_newline": b"@hi\nNN\n+\n++",
    "header_error_first": b"hi\nNN\n+\n++\n",
    "sep_error": b"@hi\nNN\n-\n++\n",
    "sep_empty_line": b"@hi\nNN\n\n++\n",
    "length_error_second": b"@a\nAC\n+\n!!\n@b\nACG\n+\n!!\n@c\nA\n+\n!\n",
    "lowercase_and_other": b"@a\nacgtnXYZ.*\n+\n!!!!!!!!!!\n",
    "nonascii": b"@a\nAC\xc3\xa9GT\n+\n\x80\xff!!~\x7f\n@b\nAC\n+\n\xfe\x01\n",
    "tabs_nul": b"@a\tb\x00\nA\x00C\n+\n\x00\t!\n",
    "only_newlines": b"\n" * 100,
    "many_blank_after": b"@a\nAC\n+\n!!\n" + b"\n" * 9000,
}


@pytest.mark.parametrize("name", sorted(EDGE))
def test_edge_case(name, torch, oracle, eng):
    check_device_vs_oracle(torch, oracle, eng, EDGE[name])


def test_edge_case_small_P(torch, oracle, eng16):
    # reads longer than P: clip counters and the len_hist overflow bin
    data = b"@a\n" + b"ACGTN" * 8 + b"\n+\n" + b"I" * 40 + b"\n@b\nACG\n+\n!!!\n"
    out, st = check_device_vs_oracle(torch, oracle, eng16, data)
    assert st.clip_seq == 24 and st.clip_qual == 24


def _rec(i, L, crlf=False):
    rng = np.random.default_rng(i)
    seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), L))
    qual = bytes(rng.integers(33, 75, L, dtype=np.uint8))
    e = b"\r\n" if crlf else b"\n"
    return b"@r%d" % i + e + seq + e + b"+" + e + qual + e


def test_long_records_beyond_halo(torch, oracle, eng, eng300):
    # records longer than the 1 KiB halo take the global-memory path; positions >= P are clipped
    data = b"".join(_rec(i, L) for i, L in enumerate([10, 1500, 150, 5000, 33000, 2, 20000, 150, 150]))
    check_device_vs_oracle(torch, oracle, eng, data)
    check_device_vs_oracle(torch, oracle, eng300, data)


def test_record_at_size_limit(torch, oracle, eng):
    # exactly BUFSIZE bytes at stream start is accepted (test `bufflen`, src/lib.rs:752-774)
    ok = b"@" + b"a" * (68 * 1024 - 8) + b"\nA\n+\nB\n"
    assert len(ok) == 68 * 1024
    check_device_vs_oracle(torch, oracle, eng, ok + _rec(1, 20))
    # clearly too long
    bad = _rec(0, 30) + b"@" + b"longid" * (68 * 1024) + b"\nA\n+\nB\n"
    out, _ = check_device_vs_oracle(torch, oracle, eng, bad)
    assert out.status == 4 and out.n_records == 1
    # long sequence line that never ends
    bad2 = _rec(0, 30) + b"@x\n" + b"A" * 100000
    out, _ = check_device_vs_oracle(torch, oracle, eng, bad2)
    assert out.status == 4


def _rec_total(n: int) -> bytes:
    """a valid record of exactly n >= 8 bytes (all of its length in the id line)"""
    return b"@" + b"a" * (n - 8) + b"\nA\n+\nB\n"


TOO_LONG_FIRST = [[], [8], [9], [15], [16], [17], [31], [45], [8, 9], [1000, 13], [69000], [40000, 40000, 11]]


@pytest.mark.parametrize("first", TOO_LONG_FIRST, ids=lambda f: "+".join(map(str, f)) or "start")
def test_too_long_band(first, torch, oracle, eng, fq):
    """Records of 69 610 .. 69 640 bytes: whether the reference accepts one depends on where Buffer::clean parks
    it (src/buffer.rs:51-72, src/lib.rs:276-283), i.e. on the stream offset mod 16 of its first byte.  Expected
    values come from the oracle's restatement of Buffer; every length goes through fqb_parse_device,
    fqb_parse_host and the FQB_F_PARTIAL refill path."""
    pre = b"".join(_rec_total(n) for n in first)
    for L in list(range(69610, 69641)):
        data = pre + _rec_total(L) + _rec_total(20)
        ores, _ = oracle.each_index(data)
        fits = (len(pre) & 15) + L <= 68 * 1024
        assert (ores.status, ores.n_records) == ((0, len(first) + 2) if fits else (4, len(first))), (first, L)
        t = to_dev(torch, data)
        eng.parse_device(t, n_own=len(data), n_avail=len(data), hist=False)
        out, _ = eng.fetch(want_stats=False)
        assert (out.status, out.n_records) == (ores.status, ores.n_records), (first, L, out)
        if ores.status:
            assert out.err_offset == ores.err_offset
        if L % 5 == 0 or not fits:
            hout, _, _ = eng.parse_host(data, hist=False, want_stats=False)
            assert (hout.status, hout.n_records) == (ores.status, ores.n_records), (first, L, hout)
            seen = []
            try:
                fq.Parser(io.BytesIO(data), engine=eng, chunk_bytes=50000).each(lambda r: seen.append(r.offset) or True)
                assert ores.status == 0
            except fq.FastqError as e:
                assert (e.status, e.offset, e.n_delivered) == (ores.status, ores.err_offset, ores.n_records)
            assert len(seen) == ores.n_records, (first, L)
    # a record that never ends: too long iff the stream holds the whole window from its start, else truncated
    for avail in (69610, 69616, 69617, 69620, 69625, 69631, 69632, 69633, 69640):
        data = pre + b"@" + b"a" * (avail - 1)
        ores, _ = oracle.each_index(data)
        assert ores.status == (4 if avail >= 68 * 1024 - (len(pre) & 15) else 5)
        t = to_dev(torch, data)
        eng.parse_device(t, n_own=len(data), n_avail=len(data), hist=False)
        out, _ = eng.fetch(want_stats=False)
        assert (out.status, out.n_records, out.err_offset) == (ores.status, ores.n_records, ores.err_offset), (first, avail)
        hout, _, _ = eng.parse_host(data, hist=False, want_stats=False)
        assert (hout.status, hout.n_records, hout.err_offset) == (ores.status, ores.n_records, ores.err_offset)


def test_too_long_band_in_a_later_shard(torch, oracle, eng):
    """the rule uses the STREAM offset of the record: a shard that starts at a nonzero stream offset"""
    for shift in (0, 5, 16, 27):
        pre = _rec_total(32 + shift)
        for L in (69616, 69617, 69620, 69627, 69632):
            data = pre + _rec_total(L) + _rec_total(20)
            ores, _ = oracle.each_index(data)
            body = data[len(pre):]
            t = to_dev(torch, body)
            eng.parse_device(t, n_own=len(body), n_avail=len(body), hist=False, stream_offset=len(pre), line_base=4)
            out, _ = eng.fetch(want_stats=False)
            assert (out.status, out.n_records + 1) == (ores.status, ores.n_records), (shift, L)   # (+ the record in front)
            if ores.status:
                assert out.err_offset == ores.err_offset


def test_histogram_drain_with_uneven_warps(torch, oracle, eng):
    """The u16 counter halves of the speculative kernel are drained by whichever warp pushes the CTA's record
    count over the mark -- independent of the other warps still being at work.  Here one warp range in 32 (the
    range whose slice of the table holds ('A', 0) and ('I', 0) under a per-warp drain) holds 1 KiB records and is
    done early, while the other 31 ranges of its CTA keep bumping those two counters 8192 times each: > 250 000
    bumps per CTA on one counter word."""
    blk = 65536
    small = b"@\nA\n+\nI\n" * (blk // 8)
    rng = np.random.default_rng(5)
    long_recs = []
    for i in range(blk // 1024):
        seq = bytes(rng.choice(np.frombuffer(b"ACGT", dtype=np.uint8), 508))
        qual = bytes(rng.integers(40, 70, 508, dtype=np.uint8))
        long_recs.append(b"@hh\n" + seq + b"\n+\n" + qual + b"\n")
    long_blk = b"".join(long_recs)
    assert len(small) == blk and len(long_blk) == blk
    n_ranges = 148 * 32
    cta = torch.cat([torch.from_numpy(np.frombuffer(long_blk if w == 3 else small, dtype=np.uint8).copy()) for w in range(32)]).cuda()
    t = torch.cat([cta.repeat(n_ranges // 32), torch.zeros(64, dtype=torch.uint8, device="cuda")])
    n = blk * n_ranges
    eng.parse_device(t, n_own=n, n_avail=n, hist=True)
    out, st = eng.fetch()
    n_small, n_long = (n_ranges - n_ranges // 32) * (blk // 8), (n_ranges // 32) * (blk // 1024)
    assert out.status == 0 and out.n_records == n_small + n_long
    assert not eng.last_path()["exact"]
    _, s_small = oracle.each_stats(small, 150)
    _, s_long = oracle.each_stats(long_blk, 150)
    k_small, k_long = n_ranges - n_ranges // 32, n_ranges // 32
    np.testing.assert_array_equal(st.qual_hist, s_small.qual_hist * k_small + s_long.qual_hist * k_long)
    np.testing.assert_array_equal(st.base_hist, s_small.base_hist * k_small + s_long.base_hist * k_long)
    np.testing.assert_array_equal(st.len_hist, s_small.len_hist * k_small + s_long.len_hist * k_long)
    assert st.n_bases == s_small.n_bases * k_small + s_long.n_bases * k_long
    assert st.clip_seq == s_long.clip_seq * k_long and st.clip_qual == s_long.clip_qual * k_long


def test_dense_newlines_list_overflow(torch, oracle, eng):
    # > LIST_CAP newlines in one 16 KiB tile: 6-byte records (empty seq/qual) and 8-byte records
    data = b"@\n\n+\n\n" * 6000 + b"@a\nA\n+\n!\n" * 3000 + b"@\n\n+\n\n" * 100
    check_device_vs_oracle(torch, oracle, eng, data)
    bad = b"@\n\n+\n\n" * 5000 + b"X\n\n+\n\n" + b"@\n\n+\n\n" * 100
    out, _ = check_device_vs_oracle(torch, oracle, eng, bad)
    assert out.status == 1 and out.n_records == 5000


# --------------------------------------------------------------------------------------------
# synthetic inputs A (fixed 150 bp) and B (variable 50..300 bp), device generator == oracle generator
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("n_rec,off", [(1, 0), (7, 0), (1000, 0), (31152, 0), (500, 321 * 12345678901 + 17)])
def test_synth_fixed_generator_and_parse(n_rec, off, torch, oracle, eng):
    n = n_rec * 321
    t = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    eng.synth_fixed(t, n, byte_off=off)
    torch.cuda.synchronize()
    got = t[:n].cpu().numpy()
    want = oracle.synth_fixed(n, 150, off)
    np.testing.assert_array_equal(got, want)
    if off % 321 == 0:
        check_device_vs_oracle(torch, oracle, eng, want.tobytes())


def test_synth_fixed_unaligned_window(torch, oracle, eng):
    # a window that starts and ends mid-record
    n, off = 100003, 777
    t = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    eng.synth_fixed(t, n, byte_off=off)
    np.testing.assert_array_equal(t[:n].cpu().numpy(), oracle.synth_fixed(n, 150, off))


@pytest.mark.parametrize("n_rec,first", [(1, 0), (2000, 0), (30000, 987654321)])
def test_synth_var_generator_and_parse(n_rec, first, torch, oracle, eng300):
    t, total = eng300.synth_var(n_rec, first=first, pad=64)
    want = oracle.synth_var(n_rec, first)
    assert total == want.size
    np.testing.assert_array_equal(t[:total].cpu().numpy(), want)
    check_device_vs_oracle(torch, oracle, eng300, want.tobytes())


def test_var_length_with_small_P(torch, oracle, eng):
    # variable 50..300 bp reads against P = 150: clipping + len_hist overflow bin
    want = oracle.synth_var(5000, 0)
    check_device_vs_oracle(torch, oracle, eng, want.tobytes())


def test_crlf_synthetic(torch, oracle, eng):
    data = b"".join(_rec(i, 150, crlf=True) for i in range(400))
    check_device_vs_oracle(torch, oracle, eng, data)


def test_count_lines(torch, oracle, eng):
    for n in (0, 1, 15, 16, 17, 321 * 1000 + 5):
        data = oracle.synth_fixed(n, 150, 0).tobytes()
        t = to_dev(torch, data)
        assert eng.count_lines(t, n) == data.count(b"\n")


# --------------------------------------------------------------------------------------------
# error in the middle of a large input: everything before the first bad record is delivered
# --------------------------------------------------------------------------------------------
@pytest.mark.parametrize("kind", ["header", "sep", "length", "truncated"])
def test_error_mid_stream(kind, torch, oracle, eng):
    n_rec = 4000
    a = bytearray(oracle.synth_fixed_records(n_rec).tobytes())
    k = 2500
    base = k * 321
    if kind == "header":
        a[base] = ord("X")
    elif kind == "sep":
        a[base + 17 + 151] = ord("-")
    elif kind == "length":
        a[base + 17 + 151 + 2 + 10] = ord("\n")   # splits the quality line
    elif kind == "truncated":
        a = a[:base + 200]
    out, st = check_device_vs_oracle(torch, oracle, eng, bytes(a))
    assert out.n_records == k and out.status != 0


# --------------------------------------------------------------------------------------------
# the predicting delimiter (fq_stream.cu: windows whose records keep the shape of the last scanned
# record are verified instead of scanned): everything that must end or defeat a prediction
# --------------------------------------------------------------------------------------------
def _plain_rec(head: bytes, L: int, i: int, eol=b"\n"):
    seq = bytes(b"ACGT"[(i * 7 + j * 3) & 3] for j in range(L))
    qual = bytes(33 + ((i + j * 5) % 41) for j in range(L))
    return b"@" + head + eol + seq + eol + b"+" + eol + qual + eol


def test_prediction_header_length_changes(torch, oracle, eng):
    # unpadded ids: the header grows by one byte at 10, 100, 1000, 10000 -- each time the prediction
    # must stop exactly there and a scan take over
    data = b"".join(_plain_rec(b"r%d" % i, 100, i) for i in range(12000))
    check_device_vs_oracle(torch, oracle, eng, data)


def test_prediction_alternating_lengths(torch, oracle, eng):
    # read lengths that change every few records (trimmed reads): predictions fail all the time
    data = b"".join(_plain_rec(b"x%05d" % i, 60 + 13 * ((i // 3) % 7), i) for i in range(9000))
    check_device_vs_oracle(torch, oracle, eng, data)


def test_prediction_crlf_and_mixed_endings(torch, oracle, eng):
    # CRLF records keep their shape (prediction on), then a block of LF records, then records whose
    # quality line alone carries the \r (raw lengths still equal? no: that is a length error) -> stop
    a = b"".join(_plain_rec(b"c%06d" % i, 120, i, eol=b"\r\n") for i in range(3000))
    b = b"".join(_plain_rec(b"l%06d" % i, 120, i) for i in range(3000))
    check_device_vs_oracle(torch, oracle, eng, a + b + a)


def _illumina(n, L=150, seed=0, crlf=False, sep_id=False, hdr_extra=0):
    """fixed-length reads whose id lines vary in length (tile / x / y coordinates)"""
    rng = np.random.default_rng(seed)
    e = b"\r\n" if crlf else b"\n"
    out = []
    for i in range(n):
        h = b"@A00123:45:HXXXXDSXX:1:%d:%d:%d 1:N:0:ATCACGTT" % (rng.integers(1101, 2678), rng.integers(1, 30000),
                                                                 rng.integers(1, 100000))
        h += b"x" * int(rng.integers(0, hdr_extra + 1))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), L))
        qual = bytes(rng.integers(33, 75, L, dtype=np.uint8))
        out.append(h + e + seq + e + b"+" + (h[1:] if sep_id else b"") + e + qual + e)
    return out


@pytest.mark.parametrize("variant", ["plain", "crlf", "L100", "long_headers", "sep_id", "L0", "L25"])
def test_prediction_varying_header_lengths(variant, torch, oracle, eng):
    """Headers of varying length (real instrument ids): the kernel searches every header end and predicts
    the rest of the record; result identical to the oracle, index included."""
    recs = {"plain": lambda: _illumina(6000), "crlf": lambda: _illumina(6000, crlf=True),
            "L100": lambda: _illumina(6000, L=100), "long_headers": lambda: _illumina(6000, hdr_extra=90),
            "sep_id": lambda: _illumina(3000, L=60, sep_id=True), "L0": lambda: _illumina(3000, L=0),
            "L25": lambda: _illumina(9000, L=25)}[variant]()   # more than 32 records per window
    check_device_vs_oracle(torch, oracle, eng, b"".join(recs))


@pytest.mark.parametrize("where", ["seq_nl", "qual_nl", "plus", "at", "short_read", "long_read", "header_200", "qual_short"])
def test_varying_headers_with_a_bad_or_odd_record(where, torch, oracle, eng):
    recs = _illumina(4000, seed=3)
    k = 2777
    r = bytearray(recs[k])
    h = r.index(b"\n")
    if where == "seq_nl":
        r[h + 40] = 10
    elif where == "qual_nl":
        r[h + 1 + 151 + 2 + 70] = 10
    elif where == "plus":
        r[h + 1 + 151] = ord("-")
    elif where == "at":
        r[0] = ord("A")
    elif where == "short_read":                       # a valid record with a shorter read: not an error
        r = bytearray(_rec(k, 90))
    elif where == "long_read":
        r = bytearray(_rec(k, 151))
    elif where == "header_200":                       # header longer than the search window: valid
        r = bytearray(b"@" + b"h" * 200 + bytes(r[h:]))
    elif where == "qual_short":                       # length mismatch
        del r[-2]
    recs[k] = bytes(r)
    check_device_vs_oracle(torch, oracle, eng, b"".join(recs))


@pytest.mark.parametrize("where", ["seq", "qual", "header", "sep"])
def test_prediction_stray_newline(where, torch, oracle, eng):
    """A '\n' inside a line of a record deep inside a predicted run: all predicted line ends are
    still '\n', so only the extra checks (header / separator scan, the empty '\n' row of the
    histogram) can notice that the lines are not what the prediction takes them for."""
    recs = [_plain_rec(b"id%07d+x" % i, 150, i) for i in range(6000)]
    k = 4321
    r = bytearray(recs[k])
    hl = len(b"id%07d+x" % k) + 2
    pos = {"header": 5, "seq": hl + 77, "sep": hl + 151, "qual": hl + 151 + 2 + 40}[where]
    if where == "sep":
        r = bytearray(recs[k].replace(b"\n+\n", b"\n+ab\n"))   # a separator line with text ...
        recs = [x.replace(b"\n+\n", b"\n+ab\n") for x in recs]  # ... in every record, so the shape predicts
        r[pos + 2] = ord("\n")
    else:
        r[pos] = ord("\n")
    recs[k] = bytes(r)
    out, _ = check_device_vs_oracle(torch, oracle, eng, b"".join(recs))
    assert out.status != 0 and out.n_records == k


def test_prediction_high_bytes(torch, oracle, eng):
    # a byte >= 0x80 in a quality line deep inside a predicted run (must not reach the dp4a addressing)
    recs = [_plain_rec(b"h%06d" % i, 150, i) for i in range(5000)]
    r = bytearray(recs[3777])
    r[len(r) - 20] = 0xC3
    recs[3777] = bytes(r)
    check_device_vs_oracle(torch, oracle, eng, b"".join(recs))


def check_count_mode(torch, oracle, engine, data: bytes):
    """delimit + index WITHOUT histograms (the predicting delimiter then verifies a window by its
    '\n' count instead of the histogram's '\n' row) == oracle.each on the same bytes."""
    t = to_dev(torch, data)
    idx = torch.zeros(len(data) + 8, dtype=torch.int32, device="cuda")
    for index in (idx, None):
        engine.parse_device(t, n_own=len(data), n_avail=len(data), hist=False, index=index)
        out, _ = engine.fetch(want_stats=False)
        ores, orecs = oracle.each(data)
        assert (out.status, out.n_records) == (ores.status, ores.n_records), (out, ores)
        if ores.status != 0:
            assert out.err_offset == ores.err_offset
        assert out.n_lines == data.count(b"\n")
    _, oidx = oracle.each_index(data)
    n = out.n_records
    got = idx[:4 * n].cpu().numpy().view(np.uint32).astype(np.uint64).reshape(n, 4)
    np.testing.assert_array_equal(got, oidx[:, 1:5])
    return out


@pytest.mark.parametrize("where", ["none", "seq", "qual", "header", "sep", "at", "plus"])
def test_count_mode_prediction(where, torch, oracle, eng):
    recs = [_plain_rec(b"id%07d" % i, 150, i) for i in range(6000)]
    k = 4321
    r = bytearray(recs[k])
    hl = len(b"id%07d" % k) + 2
    if where == "seq":
        r[hl + 77] = ord("\n")
    elif where == "qual":
        r[hl + 151 + 2 + 40] = ord("\n")
    elif where == "header":
        r[5] = ord("\n")
    elif where == "sep":
        recs = [x.replace(b"\n+\n", b"\n+ab\n") for x in recs]
        r = bytearray(recs[k])
        r[hl + 151 + 2] = ord("\n")
    elif where == "at":
        r[0] = ord("A")
    elif where == "plus":
        r[hl + 151] = ord("-")
    recs[k] = bytes(r)
    out = check_count_mode(torch, oracle, eng, b"".join(recs))
    assert (out.status != 0 and out.n_records == k) if where != "none" else out.n_records == 6000


def test_count_mode_with_non_ascii_bytes(torch, oracle, eng):
    """Without histograms nothing depends on the byte values beyond '\\n', '@', '+': ids, reads and qualities
    with bytes >= 0x80 are delimited like any others (fixed and varying shapes)."""
    rng = np.random.default_rng(11)
    recs = []
    for i in range(5000):
        L = 150 if i < 3000 else int(rng.integers(1, 200))
        head = ("@r%d é\u00fc\u4e2d" % i).encode("utf-8")
        seq = bytes(rng.integers(128, 256, L, dtype=np.uint8))
        qual = bytes(rng.integers(128, 256, L, dtype=np.uint8))
        recs.append(head + b"\n" + seq + b"\n+\n" + qual + b"\n")
    data = b"".join(recs)
    check_count_mode(torch, oracle, eng, data)
    check_device_vs_oracle(torch, oracle, eng, data)          # with histograms: the exact path counts rows >= 128


def test_clean_inputs_stay_on_the_fast_path(torch, oracle, eng, eng300):
    """Guard against silent fallbacks (parity would still hold, throughput would not): clean inputs of every
    BASELINE shape are served by the speculative kernel, fixed shapes almost entirely by predicted windows."""
    def path(engine, t, n, hist, n_rec):
        idx = torch.zeros(4 * n_rec + 8, dtype=torch.int32, device="cuda")
        engine.parse_device(t, n_own=n, n_avail=n, hist=hist, index=idx)
        out, _ = engine.fetch(want_stats=False)
        assert out.status == 0 and out.n_records == n_rec
        return engine.last_path()

    def tiled(block: bytes, reps: int):
        a = torch.from_numpy(np.frombuffer(block, dtype=np.uint8).copy()).cuda().repeat(reps)
        return torch.cat([a, torch.zeros(64, dtype=torch.uint8, device="cuda")]), len(block) * reps

    n_rec = 1600000                                     # 514 MB of fixed 150 bp records
    t = torch.zeros(n_rec * 321 + 64, dtype=torch.uint8, device="cuda")
    eng.synth_fixed(t, n_rec * 321)
    for hist in (True, False):
        p = path(eng, t, n_rec * 321, hist, n_rec)
        assert not p["exact"] and p["predicted"] > 5 * p["scanned"], p
    t, n = tiled(oracle.synth_var(20000).tobytes(), 20)            # variable 50-300 bp
    for hist in (True, False):
        p = path(eng300, t, n, hist, 20000 * 20)
        assert not p["exact"] and p["scanned"] > 0, p
    t, n = tiled(b"".join(_illumina(20000)), 30)                   # id lines of varying length
    p = path(eng, t, n, True, 20000 * 30)
    assert not p["exact"] and p["predicted"] > 5 * p["scanned"], p
    assert not path(eng, t, n, False, 20000 * 30)["exact"]
    utf8 = b"".join(("@r%d \u00e9\u4e2d" % i).encode() + b"\n" + _rec(i, 150)[_rec(i, 150).index(b"\n") + 1:] for i in range(5000))
    t, n = tiled(utf8, 20)
    assert not path(eng, t, n, False, 5000 * 20)["exact"]          # no histograms: byte values do not matter
    # an error at the very end of the stream (truncated download, trailing blank lines or garbage) is found and
    # classified without the exact pass: every record in front of it has been counted, nothing behind it
    good = oracle.synth_fixed_records(100000).tobytes()
    for tail in (good[:-1], good[:-200], good + b"\n", good + b"\n" * 50, good + b"@x\nAC\n+\n!\n" + good[:3210],
                 good + b"garbage without a newline"):
        ores, ost = oracle.each_stats(tail, 150)
        assert ores.status != 0
        d = to_dev(torch, tail)
        idx = torch.zeros(len(tail) + 8, dtype=torch.int32, device="cuda")
        for hist in (True, False):
            eng.parse_device(d, n_own=len(tail), n_avail=len(tail), hist=hist, index=idx)
            out, st = eng.fetch(want_stats=hist)
            assert (out.status, out.n_records, out.err_offset) == (ores.status, ores.n_records, ores.err_offset)
            assert out.n_lines == tail.count(b"\n") and not out.finished
            assert not eng.last_path()["exact"], (len(tail) - len(good), hist)
            if hist:
                assert_stats_equal(st, ost)


def _var_recs(n, seed=0, lo=50, hi=300, head=b"r%d"):
    rng = np.random.default_rng(seed)
    out = []
    for i in range(n):
        L = int(rng.integers(lo, hi + 1))
        seq = bytes(rng.choice(np.frombuffer(b"ACGTN", dtype=np.uint8), L))
        qual = bytes(rng.integers(33, 127, L, dtype=np.uint8))
        out.append(b"@" + (head % i) + b"\n" + seq + b"\n+\n" + qual + b"\n")
    return out


@pytest.mark.parametrize("what", ["clean", "tab_in_seq", "cr_in_qual", "nul_in_qual", "high_in_seq", "high_in_header",
                                  "ctrl_in_header", "lengths_64k", "crlf", "clip"])
def test_variable_length_variant(what, torch, oracle, eng, eng300):
    """The speculative kernel's variant for reads of varying length (validate_block / line_steps; for P <= 320 its
    table has no rows for bytes below 32): odd bytes in the sequence / quality lines must come out exact (through
    the exact path), odd bytes in id lines must not matter, lengths at the 64-position step boundaries, '\r'
    trimmed from one line only, reads longer than the tracked positions."""
    recs = _var_recs(3000, seed=7)
    if what == "lengths_64k":
        recs = [_rec(i, L) for i, L in enumerate([0, 1, 3, 4, 63, 64, 65, 127, 128, 129, 191, 192, 193, 255, 256, 257, 299, 300] * 60)]
    elif what == "crlf":
        recs = [_rec(i, L, crlf=(i % 3 == 0)) for i, L in enumerate([64, 65, 128, 50, 300, 129, 192] * 200)]
        # '\r' in front of the '\n' of ONE of the two lines only: seq() and qual() differ in length by one
        recs[5] = b"@x\n" + b"A" * 63 + b"\r\n+\n" + b"I" * 64 + b"\n"
        recs[9] = b"@y\n" + b"C" * 128 + b"\n+\n" + b"I" * 127 + b"\r\n"
    elif what == "clip":
        recs = _var_recs(1500, seed=3, lo=250, hi=700)
    k = 1234
    def poke(line, byte):
        parts = recs[k].split(b"\n")
        parts[line] = parts[line][:20] + bytes([byte]) + parts[line][21:]
        recs[k] = b"\n".join(parts)
    if what == "tab_in_seq":
        poke(1, 9)
    elif what == "cr_in_qual":
        poke(3, 13)
    elif what == "nul_in_qual":
        poke(3, 0)
    elif what == "high_in_seq":
        poke(1, 0xC3)
    elif what == "high_in_header":
        recs = [r.replace(b"@r", "@\u00e9\u4e2d".encode(), 1) for r in recs]
    elif what == "ctrl_in_header":
        recs = [r.replace(b"@r", b"@\t\x01", 1) for r in recs]
    data = b"".join(recs)
    ores, oidx = oracle.each_index(data)
    for engine in (eng300, eng):
        check_device_vs_oracle(torch, oracle, engine, data)
        p = engine.last_path()
        if what in ("clean", "high_in_header", "ctrl_in_header", "lengths_64k", "crlf", "clip"):
            assert not p["exact"], (what, engine.max_len, p)          # served by the fast path
        # the same variant without histograms (delimit + line-end index; any byte may stand anywhere in a line)
        t = to_dev(torch, data)
        idx = torch.zeros(len(data) + 8, dtype=torch.int32, device="cuda")
        engine.parse_device(t, n_own=len(data), n_avail=len(data), hist=False, index=idx)
        out, _ = engine.fetch(want_stats=False)
        assert (out.status, out.n_records, out.n_lines) == (ores.status, ores.n_records, data.count(b"\n"))
        got = idx[:4 * out.n_records].cpu().numpy().view(np.uint32).astype(np.uint64).reshape(-1, 4)
        np.testing.assert_array_equal(got, oidx[:, 1:5])
        assert not engine.last_path()["exact"], (what, engine.max_len)


def test_mid_stream_error_costs_at_most_two_parses(torch, oracle, eng):
    """A bad record in the middle of a big shard: the speculative kernel stops its range there, the ranges in front
    are verified, and fqb_fetch parses the bytes in front of the bad record once more -- no exact pass, at most about
    twice the time of a clean parse (it was ~7x through the exact path), results as the oracle's each()."""
    n_rec = 1600000                                     # 514 MB of fixed 150 bp records
    n = n_rec * 321
    t = torch.zeros(n + 64, dtype=torch.uint8, device="cuda")
    eng.synth_fixed(t, n)
    idx = torch.zeros(4 * n_rec + 8, dtype=torch.int32, device="cuda")

    def timed(hist):
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            eng.parse_device(t, n_own=n, n_avail=n, hist=hist, index=idx)
            out, st = eng.fetch(want_stats=hist)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            best = ms if best is None else min(best, ms)
        return best, out, st

    for hist in (True, False):
        clean_ms, out, _ = timed(hist)
        assert out.status == 0 and out.n_records == n_rec
        for where, kind in ((0.6, "at"), (0.97, "plus"), (0.3, "len")):
            k = int(n_rec * where)
            off = {"at": 321 * k, "plus": 321 * k + 17 + 151, "len": 321 * k + 17 + 151 + 2 + 150}[kind]
            saved = t[off:off + 1].clone()
            t[off] = {"at": ord("X"), "plus": ord("-"), "len": ord("A")}[kind]
            err_ms, out, st = timed(hist)
            p = eng.last_path()
            assert (out.status, out.n_records, out.err_offset) == ({"at": 1, "plus": 2, "len": 3}[kind], k, 321 * k), (kind, out)
            assert out.n_lines == 4 * n_rec - (1 if kind == "len" else 0) and not out.finished
            assert p["retried"] and not p["exact"], p
            assert err_ms < 2.0 * clean_ms + 0.6, (kind, hist, err_ms, clean_ms)
            if hist:
                ores, ost = oracle.each_stats(oracle.synth_fixed_records(2000, first=k - 2000).tobytes(), 150)
                assert st.n_records == k and int(st.qual_hist.sum()) == 150 * k
                got = idx[4 * (k - 1):4 * k].cpu().numpy().view(np.uint32)
                assert int(got[3]) == (321 * k - 1) & 0xFFFFFFFF
            t[off:off + 1] = saved
    # small inputs, every byte position of a bad '@' in a 40-record file: same result as the oracle
    base = oracle.synth_fixed_records(40).tobytes()
    for k in range(0, 40, 3):
        data = bytearray(base)
        data[321 * k] = ord("#")
        check_device_vs_oracle(torch, oracle, eng, bytes(data))


def test_refill_that_ends_inside_a_record_stays_on_the_fast_path(torch, oracle, eng, eng300):
    """FQB_F_PARTIAL refills (and shards without a halo) normally end inside a record: that record is the tail to carry
    over (fqb_result.tail_offset), not a reason to redo the chunk on the exact path."""
    fixed = oracle.synth_fixed_records(30000).tobytes()
    var = oracle.synth_var(20000).tobytes()
    for engine, data in ((eng, fixed), (eng300, var), (eng, var)):
        ores, oidx = oracle.each_index(data)
        for cut in (len(data) - 1, len(data) - 100, len(data) * 2 // 3, 70001):
            part = data[:cut]
            n_ok = int(np.searchsorted(oidx[:, 4], cut, side="left"))       # records complete inside the part
            tail = int(oidx[n_ok, 0]) if n_ok < len(oidx) and int(oidx[n_ok, 0]) < cut else None
            for hist in (True, False):
                out, st, idx = engine.parse_host(part, hist=hist, want_index=True, want_stats=hist, partial=True, stream_offset=32)
                assert (out.status, out.n_records) == (0, n_ok), (cut, out)
                assert out.tail_offset == (None if tail is None else tail + 32)
                assert not engine.last_path()["exact"], (cut, hist, engine.max_len)
                np.testing.assert_array_equal(idx[:4 * n_ok].reshape(n_ok, 4), oidx[:n_ok, 1:5] + 32)
                if hist:
                    _, ost = oracle.each_stats(data[:int(oidx[n_ok - 1, 4]) + 1], engine.max_len)
                    assert_stats_equal(st, ost)
            # the same through fqb_parse_device: a shard that is all there is for now (no halo, not the end of the stream)
            t = to_dev(torch, part)
            engine.parse_device(t, n_own=cut, n_avail=cut, hist=True, eof=False)
            out, _ = engine.fetch()
            assert (out.status, out.n_records, out.tail_offset) == (0, n_ok, tail) and not engine.last_path()["exact"]


def test_count_mode_varying_shapes(torch, oracle, eng):
    data = b"".join(_plain_rec(b"r%d" % i, 60 + 13 * ((i // 5) % 7), i) for i in range(9000))
    check_count_mode(torch, oracle, eng, data)
    data = b"".join(_plain_rec(b"c%06d" % i, 120, i, eol=b"\r\n") for i in range(4000))
    check_count_mode(torch, oracle, eng, data)
    # long reads with a window-filling shape: one record per 2560 / 4096-byte window and longer
    data = b"".join(_plain_rec(b"L%05d" % i, L, i) for i, L in enumerate([900, 900, 900, 1900, 1900, 2040, 2040, 5000, 900] * 40))
    check_count_mode(torch, oracle, eng, data)


@pytest.mark.parametrize("seed", list(range(12)))
def test_prediction_fuzz(seed, torch, oracle, eng, eng300):
    """Runs of records with a constant shape (what the predicting delimiter feeds on) -- or a constant
    shape but for the length of the id line -- glued together
    with shape changes, CRLF blocks, text after '+', and -- in two thirds of the seeds -- one
    anomaly somewhere: the GPU result must equal the oracle's in histogram mode and in count mode."""
    rng = np.random.default_rng(1000 + seed)
    recs = []
    n_target = int(rng.integers(4000, 9000))
    i = 0
    while len(recs) < n_target:
        run = int(rng.integers(1, 900))
        L = int(rng.choice([0, 1, 3, 36, 75, 100, 150, 151, 200, 250, 299, 300]))
        hl = int(rng.integers(1, 70))
        eol = b"\r\n" if rng.random() < 0.2 else b"\n"
        sep_txt = bytes(rng.integers(65, 91, size=int(rng.integers(0, 40))).astype(np.uint8)) if rng.random() < 0.2 else b""
        jitter = int(rng.choice([0, 0, 3, 8, 100]))               # id lines of varying length inside a run
        for _ in range(run):
            head = (b"%d" % i).ljust(hl, b"x")[:max(hl, 1)] + b"y" * int(rng.integers(0, jitter + 1))
            seq = bytes(b"ACGTN"[int(x)] for x in rng.integers(0, 5, size=L))
            qual = bytes(rng.integers(33, 75, size=L).astype(np.uint8))
            recs.append(b"@" + head + eol + seq + eol + b"+" + sep_txt + eol + qual + eol)
            i += 1
    kind = seed % 3
    if kind:
        k = int(rng.integers(1, len(recs)))
        r = bytearray(recs[k])
        pos = int(rng.integers(0, len(r) - 1))
        r[pos] = {1: 10, 2: int(rng.choice([0x80, 0xFF, 13, 64, 43]))}[kind]
        recs[k] = bytes(r)
    data = b"".join(recs)
    if seed % 4 == 3:
        data = data[:-int(rng.integers(1, 200))]                # truncated tail
    engine = eng300 if seed % 2 else eng
    check_device_vs_oracle(torch, oracle, engine, data, check_index=True)
    check_count_mode(torch, oracle, engine, data)


# --------------------------------------------------------------------------------------------
# shards: cut a small stream at EVERY byte; owner = shard where the record starts
# --------------------------------------------------------------------------------------------
def test_two_shards_every_cut(torch, oracle, eng):
    data = (b"@a\nACGT\n+\n!!!!\n@bb\nNN\n+x\n##\n@c\n\n+\n\n@dddd\nACGTACGTAC\n+\n@+@+@+@+@+\n")
    P = eng.max_len
    _, ost = oracle.each_stats(data, P)
    n = len(data)
    for cut in range(0, n + 1):
        # shard 0 = [0, cut) with the rest as halo; shard 1 = [cut, n) in its own 16-B aligned buffer
        t0 = to_dev(torch, data)
        eng.parse_device(t0, n_own=cut, n_avail=n, line_start=True, eof=True)
        o0, s0 = eng.fetch()
        lines0 = data[:cut].count(b"\n")
        assert o0.n_lines == lines0
        buf1 = torch.zeros(16 + (n - cut) + 64, dtype=torch.uint8, device="cuda")
        front = data[max(0, cut - 16):cut]
        if front:
            buf1[16 - len(front):16] = torch.from_numpy(np.frombuffer(front, dtype=np.uint8).copy())
        if n - cut:
            buf1[16:16 + n - cut] = torch.from_numpy(np.frombuffer(data[cut:], dtype=np.uint8).copy())
        eng.parse_device(buf1[16:], n_own=n - cut, n_avail=n - cut, line_base=lines0,
                         stream_offset=cut, line_start=(cut == 0), front16=(cut > 0), eof=True)
        o1, s1 = eng.fetch()
        assert o0.status == 0 and o1.status == 0, (cut, o0, o1)
        assert o0.n_records + o1.n_records == ost.n_records, cut
        np.testing.assert_array_equal(s0.words + s1.words, _oracle_words(eng, ost), err_msg=f"cut={cut}")


def _oracle_words(engine, ost):
    from fastq_rs_b200 import _lib
    L = _lib.lib()
    P = engine.max_len
    w = np.zeros(L.fqb_stats_words(P), dtype=np.uint64)
    w[0], w[1], w[2], w[3] = ost.n_records, ost.n_bases, ost.clip_seq, ost.clip_qual
    lo, bo, qo = L.fqb_stats_len_hist_off(P), L.fqb_stats_base_hist_off(P), L.fqb_stats_qual_hist_off(P)
    w[lo:lo + P + 2] = ost.len_hist
    w[bo:bo + 6 * P] = ost.base_hist.reshape(-1)
    w[qo:qo + 256 * P] = ost.qual_hist.reshape(-1)
    return w


def test_shards_synthetic_unaligned_cuts(torch, oracle, eng):
    # 4 shards of a 3 MB synthetic stream, cuts mid-record; halo = head of the next shard
    n = 321 * 9000
    data = oracle.synth_fixed(n, 150, 0).tobytes()
    _, ost = oracle.each_stats(data, eng.max_len)
    cuts = [0, 700001, 1500016, 2200333, n]
    total = np.zeros_like(_oracle_words(eng, ost))
    nrec = 0
    full = to_dev(torch, data)
    for a, b in zip(cuts[:-1], cuts[1:]):
        a16 = a & ~15  # device buffers must be 16-byte aligned: shard view starts at an aligned address
        # emulate an aligned private copy: [front16 | shard | halo]
        buf = torch.zeros(16 + (n - a) + 64, dtype=torch.uint8, device="cuda")
        lo = max(0, a - 16)
        buf[16 - (a - lo):16 + (n - a)] = full[lo:n]
        halo = min(n - b, 68 * 1024)
        eng.parse_device(buf[16:], n_own=b - a, n_avail=b - a + halo, line_base=data[:a].count(b"\n"),
                         stream_offset=a, line_start=(a == 0), front16=(a > 0), eof=(b + halo == n))
        o, s = eng.fetch()
        assert o.status == 0
        nrec += o.n_records
        total += s.words
        del a16
    assert nrec == ost.n_records
    np.testing.assert_array_equal(total, _oracle_words(eng, ost))


def test_shards_inferred_start(torch, oracle, eng):
    """FQB_F_INFER_START: shards parsed without knowing the lines in front of them report the line
    phase their first record implies; it must equal the true one, and the results the oracle's."""
    n = 321 * 9000
    data = oracle.synth_fixed(n, 150, 0).tobytes()
    _, ost = oracle.each_stats(data, eng.max_len)
    cuts = [0, 700001, 2889 * 321, 1500016, 2200333, n]      # mid-header, record start, mid-sequence, mid-quality
    total = np.zeros_like(_oracle_words(eng, ost))
    nrec = 0
    full = to_dev(torch, data)
    for a, b in zip(cuts[:-1], cuts[1:]):
        buf = torch.zeros(16 + (n - a) + 64, dtype=torch.uint8, device="cuda")
        lo = max(0, a - 16)
        buf[16 - (a - lo):16 + (n - a)] = full[lo:n]
        halo = min(n - b, 68 * 1024)
        true_base = data[:a].count(b"\n")
        # without the 16 bytes in front a shard cannot know that it starts a line: it then treats its
        # first line as the tail of a line of the previous shard (the documented contract), which only
        # matters for a cut exactly at a line start -- the phase and the newline count hold either way
        for front in ([False] if a == 0 else [False, True]):
            eng.parse_device(buf[16:], n_own=b - a, n_avail=b - a + halo, line_base=0, stream_offset=a,
                             line_start=(a == 0), front16=(front and a > 0), eof=(b + halo == n),
                             infer_start=(a > 0))
            o, s = eng.fetch()
            assert o.status == 0, (a, o)
            assert o.line_phase == (true_base & 3), (a, front, o)
            assert o.n_lines == data[a:b].count(b"\n")
        nrec += o.n_records
        total += s.words
    assert nrec == ost.n_records
    np.testing.assert_array_equal(total, _oracle_words(eng, ost))


def test_inferred_start_ambiguous_is_reported(torch, fq, eng):
    """A shard whose first record cannot be inferred (here: bytes that never form two records)
    answers FQB_E_PHASE instead of guessing."""
    data = b"ACGT" * 5000
    t = to_dev(torch, data)
    eng.parse_device(t, n_own=len(data), n_avail=len(data), line_start=False, eof=True, infer_start=True)
    o, _ = eng.fetch()
    assert o.status == fq._lib.E_PHASE


def test_sharded_driver_single_rank(torch, fq, oracle, eng):
    """sharded.py with the real engine, world = 1 (the N-rank protocol itself runs under gloo in
    tests/test_sharded.py and on 2 GPUs in bench.py --gpus 2)."""
    from fastq_rs_b200.sharded import ShardedParser, ShardSpec
    data = oracle.synth_fixed(321 * 5000, 150, 0).tobytes()
    _, ost = oracle.each_stats(data, eng.max_len)
    t = to_dev(torch, data)
    sp = ShardedParser(eng, dist=None, device="cuda")
    out, st = sp.parse(ShardSpec(t, 0, len(data), 0, 0, True))
    assert (out.status, out.n_records) == (0, 5000)
    np.testing.assert_array_equal(st.words, _oracle_words(eng, ost))


# --------------------------------------------------------------------------------------------
# streaming ring: chunk boundaries, short fills, errors in later chunks
# --------------------------------------------------------------------------------------------
def test_streaming_small_slots(fq, oracle):
    e = fq.Engine(max_len=150, slot_bytes=2 * 68 * 1024, n_slots=2)
    try:
        data = oracle.synth_fixed_records(9000).tobytes()   # ~2.9 MB -> 21 chunks
        _, ost = oracle.each_stats(data, 150)
        out, st, idx = e.parse_host(data, want_index=True)
        assert out.status == 0 and out.n_records == 9000
        assert_stats_equal(st, ost)
        _, oidx = oracle.each_index(data)
        np.testing.assert_array_equal(idx.reshape(-1, 4), oidx[:, 1:5])
        # acquire/submit protocol with short, ragged fills (a reader that returns short reads)
        e.stream_begin()
        pos, k = 0, 0
        while pos < len(data):
            slot = e.stream_acquire()
            take = min(len(slot), len(data) - pos, 1 + (k * 7919) % 50000)
            np.frombuffer(slot, dtype=np.uint8)[:take] = np.frombuffer(data[pos:pos + take], dtype=np.uint8)
            e.stream_submit(take)
            pos += take
            k += 1
        out2, st2 = e.stream_finish()
        assert out2.status == 0 and out2.n_records == 9000
        assert_stats_equal(st2, ost)
        # an error in a late chunk: counts stop exactly at the bad record
        bad = bytearray(data)
        bad[7000 * 321] = ord("X")
        ores, ost_b = oracle.each_stats(bytes(bad), 150)
        out3, st3, _ = e.parse_host(bytes(bad))
        assert (out3.status, out3.n_records, out3.err_offset) == (ores.status, 7000, ores.err_offset)
        assert_stats_equal(st3, ost_b)
        # truncated at the very end
        out4, st4, _ = e.parse_host(data[:-1])
        ores4, ost4 = oracle.each_stats(data[:-1], 150)
        assert (out4.status, out4.n_records) == (ores4.status, ores4.n_records)
        assert_stats_equal(st4, ost4)
        # record that straddles a chunk boundary while being longer than the tile halo
        big = b"".join(_rec(i, L) for i, L in enumerate([60000, 150, 60000, 30000, 150, 65000, 10] * 3))
        ores5, ost5 = oracle.each_stats(big, 150)
        out5, st5, _ = e.parse_host(big)
        assert (out5.status, out5.n_records) == (ores5.status, ores5.n_records)
        assert_stats_equal(st5, ost5)
    finally:
        e.close()


def test_parse_host_pinned_index(fq, oracle):
    """A host_index in pinned memory is written by the device chunk by chunk (no synchronisation per
    chunk); same entries as the pageable path, also with an error in a late chunk and a small cap."""
    import ctypes
    from fastq_rs_b200 import _lib
    L = _lib.lib()
    e = fq.Engine(max_len=150, slot_bytes=2 * 68 * 1024, n_slots=2)
    try:
        good = oracle.synth_fixed_records(9000).tobytes()
        bad = bytearray(good)
        bad[7000 * 321] = ord("X")
        for data, cap in ((good, 36000), (bytes(bad), 36000), (good, 1001), (good[:-1], 36000), (b"", 16)):
            ores, oidx = oracle.each_index(data)
            _, _, want = e.parse_host(data, hist=False, want_index=True, want_stats=False)    # pageable buffer
            p = ctypes.c_void_p()
            assert L.fqb_host_alloc(cap * 4, ctypes.byref(p)) == 0
            res, got = _lib.Result(), ctypes.c_uint64(0)
            a = np.frombuffer(data, dtype=np.uint8)
            rc = L.fqb_parse_host(e.ctx, a.ctypes.data if a.size else None, a.size, 0, _lib.F_INDEX, ctypes.byref(res), None,
                                  p, cap, ctypes.byref(got))
            assert rc == 0 and (res.status, res.n_records) == (ores.status, ores.n_records)
            idx = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_uint32)), shape=(cap,))
            n = min(cap, want.size)
            assert got.value == n
            np.testing.assert_array_equal(idx[:n].astype(np.uint64), want[:n])
            np.testing.assert_array_equal(idx[:min(n, 4 * ores.n_records)].astype(np.uint64),
                                          oidx[:, 1:5].reshape(-1)[:min(n, 4 * ores.n_records)])
            L.fqb_host_free(p)
    finally:
        e.close()


class _ShortReader:
    """A reader that returns short, ragged reads (no readinto)."""

    def __init__(self, data: bytes, with_readinto: bool):
        self._b, self._k = io.BytesIO(data), 0
        if with_readinto:
            self.readinto = self._readinto

    def read(self, n):
        self._k += 1
        return self._b.read(min(n, 1 + (self._k * 7919) % 30011))

    def _readinto(self, mv):
        b = self.read(len(mv))
        mv[:len(b)] = b
        return len(b)


@pytest.mark.parametrize("chunk", [1, 4097, 70000, 1 << 20])
def test_each_over_a_reader_in_refills(chunk, fq, oracle, eng):
    """Bounded-memory generic-closure path: the reader is consumed `chunk` bytes at a time
    (FQB_F_PARTIAL + carry-over of the incomplete trailing record); records, order and the error
    -- kind, stream offset, records delivered before it -- equal the oracle's each()."""
    recs = [_rec(i, L) for i, L in enumerate([150, 0, 3, 150, 60000, 151, 150, 1, 30000, 150] * (1 if chunk < 4097 else 6))]
    good = b"".join(recs)
    if chunk == 1:
        good = b"".join(_rec(i, L) for i, L in enumerate([3, 0, 5]))
    cases = [good, good[:-1], good + b"\n", good[:len(good) // 2] + b"X" + good[len(good) // 2:], good + b"@tail\nAC\n"]
    if chunk > 1:
        cases.append(good + _rec(99, 40000) + _rec(100, 5))   # a record that is too long, seen across refills
    for k, data in enumerate(cases):
        ores, oidx = oracle.each_index(data)
        for with_readinto in (False, True):
            # the fast paths over the same reader (pinned ring fed by read() or readinto())
            try:
                assert fq.Parser(_ShortReader(data, with_readinto), engine=eng).count() == ores.n_records
                assert ores.status == 0
            except fq.FastqError as e:
                assert e.status == ores.status and e.n_delivered == ores.n_records
            seen, err = [], None
            try:
                fin = fq.Parser(_ShortReader(data, with_readinto), engine=eng, chunk_bytes=chunk).each(
                    lambda r: seen.append((r.offset, bytes(r.data), r.head(), r.seq(), r.qual())) or True)
                assert fin is True
            except fq.FastqError as e:
                err = e
            assert len(seen) == ores.n_records, (k, chunk)
            assert (err.status if err else 0) == ores.status, (k, chunk)
            if err:
                assert err.offset == ores.err_offset and err.n_delivered == ores.n_records
            for i, (_, raw, head, seq, qual) in enumerate(seen):
                s, e0, e1, e2, e3 = (int(x) for x in oidx[i])
                assert raw == data[s:e3 + 1]
                assert head == data[s + 1:e0] and seq == data[e0 + 1:e1] and qual == data[e2 + 1:e3]
        # record_sets / parallel_each over the same reader: same records (the batch holding the bad one is dropped)
        n_sets = 0
        try:
            for s in fq.Parser(_ShortReader(data, True), engine=eng, chunk_bytes=chunk).record_sets():
                n_sets += s.len()
            assert ores.status == 0
        except fq.FastqError as e:
            assert e.status == ores.status
        assert n_sets <= ores.n_records and (ores.status != 0 or n_sets == ores.n_records)


@pytest.fixture(scope="module")
def eng_small_slots(fq):
    e = fq.Engine(max_len=150, slot_bytes=160 * 1024, n_slots=4)      # many batches, records that straddle slots
    yield e
    e.close()


class _FailingReader:
    def __init__(self, data, fail_at):
        self._b, self._n, self._fail_at = io.BytesIO(data), 0, fail_at

    def read(self, n):
        if self._n >= self._fail_at:
            raise OSError("disk on fire")
        b = self._b.read(min(n, 50000))
        self._n += len(b)
        return b


def test_batch_mode_each_and_record_sets(fq, oracle, eng_small_slots):
    """The asynchronous generic-closure path (fqb_batch_begin / fqb_next_batch / fqb_release_batch): a reader
    thread feeds the pinned ring while the caller walks the batches already delimited.  Records, their order, the
    error and the RecordSets equal the oracle's each() / record_sets()."""
    e = eng_small_slots
    recs = [_rec(i, L) for i, L in enumerate([150, 0, 3, 150, 30000, 151, 150, 1, 34000, 150, 20000, 7] * 12)]
    good = b"".join(recs)
    cases = [good, good[:-1], good + b"\n", good[:len(good) // 2] + b"X" + good[len(good) // 2:], good + b"@tail\nAC\n",
             good + _rec(99, 40000) + _rec(100, 5), b"", b"@a\nA\n+\nI\n", good[:321 * 3]]
    for k, data in enumerate(cases):
        ores, oidx = oracle.each_index(data)
        for with_readinto in (False, True):
            seen, err = [], None
            try:
                fin = fq.Parser(_ShortReader(data, with_readinto), engine=e).each(
                    lambda r: seen.append((bytes(r.data), r.head(), r.seq(), r.qual())) or True)
                assert fin is True
            except fq.FastqError as ex:
                err = ex
            assert len(seen) == ores.n_records, (k, len(seen), ores.n_records)
            assert (err.status if err else 0) == ores.status, k
            if err:
                assert err.offset == ores.err_offset and err.n_delivered == ores.n_records
            for i, (raw, head, seq, qual) in enumerate(seen):
                s, e0, e1, e2, e3 = (int(x) for x in oidx[i])
                assert raw == data[s:e3 + 1] and head == data[s + 1:e0] and seq == data[e0 + 1:e1] and qual == data[e2 + 1:e3]
        ostatus, osets = oracle.record_sets(data)
        sets, serr = [], 0
        try:
            for s in fq.Parser(io.BytesIO(data), engine=e).record_sets():
                sets.append([bytes(r.data) for r in s.iter()])
        except fq.FastqError as ex:
            serr = ex.status
        assert serr == ostatus
        assert sets == [[r.raw for r in x] for x in osets], k
    # early stop: the closure says no, the reader thread is told to stop, the context is reusable at once
    n = []
    assert fq.Parser(io.BytesIO(good), engine=e).each(lambda r: n.append(1) or len(n) < 5) is False and len(n) == 5
    assert fq.Parser(io.BytesIO(good), engine=e).count() == len(recs)
    # a reader that fails: its error reaches the caller, after the records in front of it
    with pytest.raises(OSError):
        fq.Parser(_FailingReader(good, 300000), engine=e).each(lambda r: True)
    out = fq.Parser(io.BytesIO(good), engine=e).parallel_each(3, lambda sets: sum(s.len() for s in sets))
    assert sum(out) == len(recs)


def test_c_program_drives_the_batch_mode(oracle, tmp_path):
    """tests/c/batch_each.c: a plain C program (pthread producer + consumer loop) over the C ABI -- no Python, no
    torch in the process -- against the oracle's each(): record count, bases, a hash over every record byte, error."""
    import subprocess
    root = os.path.dirname(HERE)
    exe = str(tmp_path / "batch_each")
    subprocess.check_call(["gcc", "-O2", "-Wall", "-I", os.path.join(root, "include"), os.path.join(HERE, "c", "batch_each.c"),
                           "-o", exe, "-L", os.path.join(root, "fastq_rs_b200"), "-l:libfastq_b200.so", "-lpthread",
                           "-Wl,-rpath," + os.path.join(root, "fastq_rs_b200")])
    good = oracle.synth_fixed_records(40000).tobytes() + b"".join(_rec(i, L, crlf=(i % 2 == 0)) for i, L in enumerate([5, 30000, 0, 151] * 20))
    for k, data in enumerate((good, good[:-7], good[:5000000] + b"?" + good[5000000:], b"")):
        path = tmp_path / f"in{k}.fastq"
        path.write_bytes(data)
        out = subprocess.run([exe, str(path), "1024"], capture_output=True, text=True, timeout=120)
        assert out.returncode == 0, out.stderr
        status, n_rec, n_bases, fnv, err_off, n_batches = (int(x) for x in out.stdout.split())
        res, recs = oracle.each(data)
        bases = sum(len(r.seq) for r in recs)
        end = (recs[-1].offset + len(recs[-1].raw)) if recs else 0
        arr = np.frombuffer(data, dtype=np.uint8)[:end].astype(np.uint64)        # the delivered records, back to back
        with np.errstate(over="ignore"):
            chk = int((arr.sum() * np.uint64(1000003) + (arr * np.arange(1, end + 1, dtype=np.uint64)).sum()) & np.uint64(0xFFFFFFFFFFFFFFFF)) if end else 0
        assert (status, n_rec, n_bases, fnv) == (res.status, res.n_records, bases, chk), (k, out.stdout)
        if res.status:
            assert err_off == res.err_offset
        assert n_batches >= max(1, (res.err_offset if res.status else len(data)) >> 20), out.stderr


def test_each_zipped_over_readers(fq, oracle, eng):
    """src/lib.rs:577-609 over two readers consumed in refills of different sizes."""
    a = oracle.synth_fixed_records(3000).tobytes()
    b = oracle.synth_fixed_records(2500).tobytes()
    pairs = []
    fin = fq.each_zipped(fq.Parser(io.BytesIO(a), engine=eng, chunk_bytes=100000),
                         fq.Parser(io.BytesIO(b), engine=eng, chunk_bytes=77777),
                         lambda r1, r2: pairs.append((r1 is not None, r2 is not None)) or (True, True))
    assert fin == (True, True)
    assert sum(1 for p in pairs if p[0]) == 3000 and sum(1 for p in pairs if p[1]) == 2500
    assert pairs[:2500] == [(True, True)] * 2500


# --------------------------------------------------------------------------------------------
# record filter: validate_dna / validate_dnan as GPU predicates + compaction
# --------------------------------------------------------------------------------------------
def _filter_cases(oracle):
    rng = np.random.default_rng(7)
    recs = []
    for i in range(3000):
        L = int(rng.integers(0, 200))
        r = bytearray(_rec(i, L, crlf=(i % 11 == 0)))
        if L and i % 3 == 0:      # plant a byte outside the alphabets somewhere in the record
            pos = int(rng.integers(0, len(r)))
            if r[pos] not in b"\r\n@+":
                r[pos] = int(rng.choice(np.frombuffer(b"acgtnXN.RYKM\x00\xff", dtype=np.uint8)))
        recs.append(bytes(r))
    mixed = b"".join(recs)
    return {
        "mixed": mixed,
        "synthetic_fixed": oracle.synth_fixed_records(5000).tobytes(),
        "synthetic_var": oracle.synth_var(3000).tobytes(),
        "empty": b"",
        "one_empty_seq": b"@id\n\n+\n\n",
        "lone_cr_seq": b"@a\n\r\n+\n\r\n@b\nAC\r\n+\n!!\r\n",
        "error_behind": mixed[:20000] + b"\nGARBAGE",
        "truncated": mixed[:-3],
        "long": b"".join(_rec(i, L) for i, L in enumerate([30000, 5, 33000, 0, 1, 2, 3, 4, 150])),
    }


@pytest.mark.parametrize("mode", [0, 1, 2])
def test_filter_device(mode, fq, torch, oracle, eng):
    for name, data in _filter_cases(oracle).items():
        ores, n_kept, want = oracle.each_filter(data, mode)
        d = to_dev(torch, data)
        idx = torch.zeros(len(data) + 8, dtype=torch.int32, device="cuda")
        eng.parse_device(d, n_own=len(data), n_avail=len(data), hist=False, index=idx)
        out, _ = eng.fetch(want_stats=False)
        assert (out.status, out.n_records) == (ores.status, ores.n_records), name
        dst = torch.full((len(data) + 16,), 0xEE, dtype=torch.uint8, device="cuda")
        eng.filter_device(d, idx, out.n_records, mode, dst)
        got_n, got_b = eng.fetch_filter()
        assert (got_n, got_b) == (n_kept, len(want)), name
        h = dst.cpu().numpy()
        assert h[:got_b].tobytes() == want, name
        assert (h[got_b:] == 0xEE).all(), name           # nothing written behind the output
        # an output buffer that is too small: totals still reported, nothing written behind its end
        if got_b > 100:
            small = torch.full((got_b // 2 + 16,), 0xEE, dtype=torch.uint8, device="cuda")
            eng.filter_device(d, idx, out.n_records, mode, small[:got_b // 2])
            assert eng.fetch_filter() == (n_kept, len(want))
            hs = small.cpu().numpy()
            cap = got_b // 2
            assert (hs[cap:] == 0xEE).all()
            neq = np.nonzero(hs[:cap] != np.frombuffer(want[:cap], dtype=np.uint8))[0]
            x = int(neq[0]) if neq.size else cap          # whole records only: the rest stays untouched
            assert (hs[x:cap] == 0xEE).all() and cap - x < 69632


def test_filter_nonzero_stream_offset_and_wrap(fq, torch, oracle, eng):
    """Index entries are the low 32 bits of stream offsets: a shard whose offsets cross a 4 GiB
    boundary filters exactly like the same bytes at offset 0."""
    data = oracle.synth_fixed_records(4000).tobytes()
    _, n_kept, want = oracle.each_filter(data, 1)
    d = to_dev(torch, data)
    for off in (0, (1 << 32) - 321 * 1000 - 7, (5 << 32) - 100, (1 << 40) + 12345):
        idx = torch.zeros(len(data) + 8, dtype=torch.int32, device="cuda")
        eng.parse_device(d, n_own=len(data), n_avail=len(data), hist=False, index=idx, stream_offset=off)
        out, _ = eng.fetch(want_stats=False)
        assert out.status == 0 and out.n_records == 4000
        dst = torch.zeros(len(data), dtype=torch.uint8, device="cuda")
        eng.filter_device(d, idx, out.n_records, fq._lib.KEEP_DNA, dst, stream_offset=off)
        assert eng.fetch_filter() == (n_kept, len(want))
        assert dst[:len(want)].cpu().numpy().tobytes() == want


def test_filter_shard_that_starts_inside_a_record(fq, torch, oracle, eng):
    """A later shard of a stream: its index begins with the line ends of the record cut by the shard start;
    the filter is handed the index from the first whole record on (not 16-byte aligned) and that record's offset."""
    recs = [_rec(i, int(L)) for i, L in enumerate(np.random.default_rng(5).integers(20, 120, 3000))]
    for k in (2, 7):                                   # plant lowercase bases in a few records
        for j in range(k, 3000, 9):
            r = bytearray(recs[j]); r[r.index(b"\n") + 3] = ord("a"); recs[j] = bytes(r)
    data = b"".join(recs)
    starts = np.cumsum([0] + [len(r) for r in recs])
    for cut_rec, inside in ((100, 7), (555, 1), (1200, 30), (2000, 16)):
        c = (int(starts[cut_rec]) + inside) // 16 * 16      # shard start, inside record `cut_rec` (or at its start)
        rec0 = int(np.searchsorted(starts, c, side="left"))  # first record that starts in the shard
        shard = data[c:]
        d = to_dev(torch, shard)
        idx = torch.zeros(len(shard) + 8, dtype=torch.int32, device="cuda")
        eng.parse_device(d, n_own=len(shard), n_avail=len(shard), hist=False, index=idx, stream_offset=c,
                         line_base=data[:c].count(b"\n"), line_start=(c == 0 or data[c - 1:c] == b"\n"))
        out, _ = eng.fetch(want_stats=False)
        assert out.status == 0 and out.n_records == 3000 - rec0
        K = data[c:int(starts[rec0])].count(b"\n")          # line ends of the cut record that lie in the shard
        _, n_kept, want = oracle.each_filter(data[int(starts[rec0]):], 1)
        dst = torch.zeros(len(shard), dtype=torch.uint8, device="cuda")
        eng.filter_device(d, idx[K:], out.n_records, fq._lib.KEEP_DNA, dst, stream_offset=c, first_offset=int(starts[rec0]))
        assert eng.fetch_filter() == (n_kept, len(want))
        assert dst[:len(want)].cpu().numpy().tobytes() == want


def test_parser_filter_to(fq, oracle, eng):
    data = _filter_cases(oracle)["mixed"]
    for keep, mode in (("all", 0), ("dna", 1), ("dnan", 2)):
        _, n_kept, want = oracle.each_filter(data, mode)
        w = io.BytesIO()
        assert fq.Parser(data, engine=eng).filter_to(w, keep) == n_kept
        assert w.getvalue() == want
    bad = data[:20000] + b"\nGARBAGE"
    _, n_kept, want = oracle.each_filter(bad, 2)
    w = io.BytesIO()
    with pytest.raises(fq.FastqError):
        fq.Parser(bad, engine=eng).filter_to(w, "dnan")
    assert w.getvalue() == want


def test_fastq_count_example_on_a_10mb_file(fq, oracle, tmp_path):
    """BASELINE config 1 (examples/fastq-count.rs on a 10 MB synthetic 150 bp file: 31 152 records),
    through parse_path -> Parser.count() (reader thread -> pinned ring -> kernels) and through
    Parser.each with a counting closure; plus Parser.stats() streamed from the file."""
    import subprocess
    import sys
    n_rec = 31152
    data = oracle.synth_fixed_records(n_rec).tobytes()
    assert len(data) == 9999792
    path = tmp_path / "reads.fastq"
    path.write_bytes(data)
    assert fq.parse_path(str(path), lambda parser: parser.count()) == n_rec
    seen = []
    assert fq.parse_path(str(path), lambda parser: parser.each(lambda r: seen.append(len(r.seq())) or True)) is True
    assert len(seen) == n_rec and set(seen) == {150}
    out, st = fq.parse_path(str(path), lambda parser: parser.stats())
    _, ost = oracle.each_stats(data, 150)
    assert out.status == 0 and out.n_records == n_rec
    assert_stats_equal(st, ost)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    got = subprocess.check_output([sys.executable, os.path.join(root, "examples", "fastq_count.py"), str(path)], text=True)
    assert int(got.strip()) == n_rec
    # the other example programs of the reference (fastq-count-thread.rs, multiple-files.rs) and the stats / filter CLI
    ex = os.path.join(root, "examples")
    got = subprocess.check_output([sys.executable, os.path.join(ex, "fastq_count_thread.py"), str(path), "3"], text=True)
    assert int(got.strip()) == n_rec
    path2 = tmp_path / "reads2.fastq"
    path2.write_bytes(data[:321 * 1000])
    got = subprocess.check_output([sys.executable, os.path.join(ex, "multiple_files.py"), str(path), str(path2)], text=True)
    assert got.strip() == f"Number of reads: ({n_rec}, 1000)"
    got = subprocess.check_output([sys.executable, os.path.join(ex, "fastq_stats_filter.py"), "stats", str(path)], text=True)
    assert got.splitlines()[0].startswith(f"records {n_rec}  bases {n_rec * 150}") and len(got.splitlines()) == 151
    kept = tmp_path / "kept.fastq"
    got = subprocess.check_output([sys.executable, os.path.join(ex, "fastq_stats_filter.py"), "filter", str(path), str(kept), "dna"], text=True)
    _, n_kept, want = oracle.each_filter(data, 1)
    assert int(got.strip()) == n_kept and kept.read_bytes() == want
    # a truncated file raises the reference's error after counting nothing more
    bad = tmp_path / "bad.fastq"
    bad.write_bytes(data[:-1])
    with pytest.raises(fq.FastqError, match="truncated"):
        fq.parse_path(str(bad), lambda parser: parser.count())


# --------------------------------------------------------------------------------------------
# property test: random valid FASTQ, randomly mutated (reference fuzz targets, fuzz/fuzz_targets/*.rs)
# --------------------------------------------------------------------------------------------
def test_hypothesis_mutations(torch, oracle, eng):
    from hypothesis import given, settings, strategies as st

    @settings(max_examples=150, deadline=None)
    @given(st.integers(0, 2**32 - 1), st.integers(1, 60), st.integers(0, 4))
    def run(seed, n_rec, n_mut):
        rng = np.random.default_rng(seed)
        recs = []
        for i in range(n_rec):
            L = int(rng.integers(0, 40)) if rng.random() < 0.8 else int(rng.integers(100, 400))
            recs.append(_rec(int(rng.integers(0, 1 << 30)), L, crlf=bool(rng.random() < 0.2)))
        data = bytearray(b"".join(recs))
        for _ in range(n_mut):
            if not data:
                break
            p = int(rng.integers(0, len(data)))
            op = rng.integers(0, 4)
            if op == 0:
                data[p] = int(rng.integers(0, 256))
            elif op == 1:
                del data[p]
            elif op == 2:
                data.insert(p, int(rng.choice([10, 13, 43, 64, 65])))
            else:
                data = data[:p]
        check_device_vs_oracle(torch, oracle, eng, bytes(data))

    run()


# --------------------------------------------------------------------------------------------
# scale: 1 GiB synthetic stream, size-independent properties (count, checksums) + oracle on a prefix
# --------------------------------------------------------------------------------------------
def test_large_synthetic_properties(torch, oracle, eng):
    n_rec = 3_000_000          # ~0.96 GB
    n = n_rec * 321
    t = torch.empty(n + 64, dtype=torch.uint8, device="cuda")
    eng.synth_fixed(t, n)
    idx = torch.empty(4 * n_rec + 8, dtype=torch.int32, device="cuda")
    eng.parse_device(t, n_own=n, n_avail=n, index=idx)
    out, st = eng.fetch()
    assert out.status == 0 and out.n_records == n_rec and out.n_lines == 4 * n_rec
    assert st.n_bases == 150 * n_rec
    # every position sees every record exactly once, in both histograms
    assert (st.base_hist.sum(axis=1) == n_rec).all() and (st.qual_hist.sum(axis=1) == n_rec).all()
    assert st.len_hist[150] == n_rec
    # offsets are an arithmetic progression for fixed-length records: closed form check of all of them
    got = idx[:4 * n_rec].view(n_rec, 4).to(torch.int64) & 0xFFFFFFFF
    rec0 = torch.arange(n_rec, device="cuda", dtype=torch.int64) * 321
    for k, rel in enumerate((16, 167, 169, 320)):
        assert torch.equal(got[:, k], (rec0 + rel) & 0xFFFFFFFF)
    # oracle on the first 40 000 records must match the histogram of the same prefix
    m = 40_000
    eng.parse_device(t, n_own=m * 321, n_avail=m * 321)
    out2, st2 = eng.fetch()
    _, ost = oracle.each_stats(t[:m * 321].cpu().numpy(), 150)
    assert_stats_equal(st2, ost)
