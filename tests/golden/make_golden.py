#!/usr/bin/env python3
"""Writes tests/golden/reference_unit_tests.json.

The reference crate (aseyboldt/fastq-rs) is Rust and cannot be executed in this
environment, so these vectors are a TRANSCRIPTION of the inputs and assertions of the
reference's own unit tests (src/lib.rs:611-811) and of the one executable doc-test
(src/lib.rs:474-508).  Each entry cites the lines it was taken from.  `expect` only holds
what the reference test itself asserts (e.g. `is_err()` without a message where the test
checks no more than that); `expect_kind` adds the error kind our restatement derives from
src/lib.rs / src/records.rs and is marked as not asserted by the reference.

Inputs that are megabytes long in the reference test are stored as a generator spec
(`input_gen`): prefix + unit*times + suffix.
"""
import base64
import json
import os

BUFSIZE = 68 * 1024  # src/lib.rs:129


def b64(b: bytes) -> str:
    return base64.b64encode(b).decode()


def rec(head, seq, qual, write=None):
    d = {"head": b64(head), "seq": b64(seq), "qual": b64(qual)}
    if write is not None:
        d["write"] = b64(write)
    return d


VECTORS = [
    {
        "name": "correct", "cite": "src/lib.rs:616-652", "api": "each",
        "input": b64(b"@hi\nNN\n+\n++\n@hallo\nTCC\n+\nabc\n"),
        "expect": {"ok": True, "records": [
            rec(b"hi", b"NN", b"++", b"@hi\nNN\n+\n++\n"),
            rec(b"hallo", b"TCC", b"abc", b"@hallo\nTCC\n+\nabc\n")]},
    },
    {
        "name": "empty_id", "cite": "src/lib.rs:654-666", "api": "each",
        "input": b64(b"@\nNN\n+\n++\n"),
        "expect": {"ok": True, "records": [rec(b"", b"NN", b"++")]},
    },
    {
        "name": "missing_lines", "cite": "src/lib.rs:668-686", "api": "each",
        "input": b64(b"@hi\nNN\n+\n++\n@hi\nNN"),
        # record 1 delivered, then Err(kind == InvalidData)
        "expect": {"ok": False, "invalid_data": True, "records": [rec(b"hi", b"NN", b"++")]},
        "expect_kind": "truncated",
    },
    {
        "name": "truncated", "cite": "src/lib.rs:688-697", "api": "each",
        "input": b64(b"@hi\nNN\n+\n++"),
        "expect": {"ok": False, "records": []},  # closure never called
        "expect_kind": "truncated",
    },
    {
        "name": "second_idline", "cite": "src/lib.rs:699-714", "api": "each",
        "input": b64(b"@hi\nNN\n+hi\n++\n@hi\nNN\n+hi\n++\n"),
        "expect": {"ok": True, "records": [
            rec(b"hi", b"NN", b"++", b"@hi\nNN\n+hi\n++\n"),
            rec(b"hi", b"NN", b"++", b"@hi\nNN\n+hi\n++\n")]},
    },
    {
        "name": "windows_lineend", "cite": "src/lib.rs:716-727", "api": "each",
        "input": b64(b"@hi\r\nNN\r\n+\r\n++\r\n@hi\r\nNN\r\n+\r\n++\r\n"),
        "expect": {"ok": True, "records": [rec(b"hi", b"NN", b"++"), rec(b"hi", b"NN", b"++")]},
    },
    {
        "name": "length_mismatch", "cite": "src/lib.rs:729-738", "api": "each",
        "input": b64(b"@hi\nNN\n+\n+\n"),
        "expect": {"ok": False, "records": []},
        "expect_kind": "length",
    },
    {
        "name": "huge_incomplete", "cite": "src/lib.rs:740-750", "api": "each",
        "input_gen": {"prefix": b64(b"@"), "unit": b64(b"longid"), "times": BUFSIZE,
                      "suffix": b64(b"")},
        "expect": {"ok": False},
        "expect_kind": "too_long",
    },
    {
        "name": "bufflen", "cite": "src/lib.rs:752-774", "api": "parallel_each",
        "n_threads": 2,
        "input_gen": {"prefix": b64(b"@"), "unit": b64(b"a"), "times": BUFSIZE - 8,
                      "suffix": b64(b"\nA\n+\nB\n")},
        "expect": {"ok": True, "count": 1},
    },
    {
        "name": "refset", "cite": "src/lib.rs:776-791", "api": "record_sets",
        "input": b64(b"@hi\nNN\n+\n++\n@hi\nNN\n+\n++\n"),
        "expect": {"ok": True, "count": 2,
                   "records": [rec(b"hi", b"NN", b"++"), rec(b"hi", b"NN", b"++")]},
    },
    {
        "name": "refset_incomplete", "cite": "src/lib.rs:793-798", "api": "record_sets",
        "input": b64(b"@hi\nNN\n+\n++\n@hi\nNN\n+\n++"),
        "expect": {"ok": False},  # any(|x| x.is_err())
        "expect_kind": "truncated",
    },
    {
        "name": "refset_huge_incomplete", "cite": "src/lib.rs:800-810", "api": "record_sets",
        "input_gen": {"prefix": b64(b"@"), "unit": b64(b"longid"), "times": BUFSIZE,
                      "suffix": b64(b"")},
        "expect": {"ok": False},
        "expect_kind": "too_long",
    },
    {
        "name": "doctest_parallel_each", "cite": "src/lib.rs:474-508", "api": "parallel_each",
        "n_threads": 4,
        "input": b64(b"@hi\nATTAATTAATTA\n+\n++++++++++++\n"),
        "expect": {"ok": True, "count": 1,
                   "records": [rec(b"hi", b"ATTAATTAATTA", b"++++++++++++")]},
    },
]


def main():
    out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "reference_unit_tests.json")
    with open(out, "w") as f:
        json.dump({"source": "aseyboldt/fastq-rs @4b510b2 src/lib.rs (transcribed)",
                   "bufsize": BUFSIZE, "vectors": VECTORS}, f, indent=1)
    print("wrote", out, len(VECTORS), "vectors")


if __name__ == "__main__":
    main()
